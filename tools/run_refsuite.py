#!/usr/bin/env python
"""Runs a selection of the reference's own tests (baseline/_ref/reference_suite) on the GPU box, once per backend,
and writes per-test outcomes to gpurun_out/refsuite_<backend>.txt.

    python tools/run_refsuite.py [--backend b200|reference|both] [--select FILE] [-- extra pytest args]

The selection file holds pytest arguments (node ids / -k expressions), one per line; default tests/refsuite/selection.txt.
"""
import argparse
import os
import subprocess
import sys
import xml.etree.ElementTree as ET

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = os.path.join(ROOT, "baseline", "_ref", "reference_suite")


def run(backend: str, select, extra, out_dir: str, timeout: int):
    xml = os.path.join(out_dir, f"refsuite_{backend}.xml")
    env = dict(os.environ)
    for k in ("MASTER_PORT", "MASTER_ADDR", "RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)  # the suite's dist_init fixture picks a free port of its own (a parent test may hold one)
    env["GSB200_DROPIN"] = "1" if backend == "b200" else "0"
    env["GSB200_DROPIN_STATS"] = os.path.join(out_dir, f"refsuite_{backend}_dropin_stats.json")
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests", "refsuite"), os.path.join(ROOT, "baseline", "_ref"), env.get("PYTHONPATH", "")])
    cmd = [sys.executable, "-m", "pytest", "-p", "gsb200_refsuite_plugin", "-q", "-p", "no:cacheprovider", f"--junitxml={xml}",
           "-o", "junit_family=xunit1"] + list(select) + list(extra)
    r = subprocess.run(cmd, cwd=SUITE, env=env, capture_output=True, text=True, timeout=timeout)
    with open(os.path.join(out_dir, f"refsuite_{backend}.log"), "w") as f:
        f.write(r.stdout[-200000:] + "\n---- stderr ----\n" + r.stderr[-20000:])
    rows = []
    if os.path.exists(xml):
        for tc in ET.parse(xml).getroot().iter("testcase"):
            name = f"{tc.get('classname', '')}::{tc.get('name', '')}"
            if tc.find("failure") is not None or tc.find("error") is not None:
                node = tc.find("failure") if tc.find("failure") is not None else tc.find("error")
                msg = (node.get("message") or "").replace("\n", " ")[:200]
                rows.append(("FAILED", name, msg))
            elif tc.find("skipped") is not None:
                msg = (tc.find("skipped").get("message") or "").replace("\n", " ")[:160]
                kind = "XFAIL" if tc.find("skipped").get("type") == "pytest.xfail" else "SKIPPED"
                rows.append((kind, name, msg))
            else:
                rows.append(("PASSED", name, ""))
    with open(os.path.join(out_dir, f"refsuite_{backend}.txt"), "w") as f:
        counts = {}
        for k, _, _ in rows:
            counts[k] = counts.get(k, 0) + 1
        f.write(f"# backend={backend} rc={r.returncode} " + " ".join(f"{k}={v}" for k, v in sorted(counts.items())) + "\n")
        for k, n, m in rows:
            f.write(f"{k:8s} {n}" + (f"   # {m}" if m else "") + "\n")
    print(f"[refsuite] backend={backend} rc={r.returncode} {counts}")
    return r.returncode, rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="both")
    ap.add_argument("--select", default=os.path.join(ROOT, "tests", "refsuite", "selection.txt"))
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    ap.add_argument("--timeout", type=int, default=3000)
    ap.add_argument("extra", nargs="*")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    select = [ln.strip() for ln in open(a.select) if ln.strip() and not ln.startswith("#")]
    rc = 0
    for b in (["reference", "b200"] if a.backend == "both" else [a.backend]):
        rc |= run(b, select, a.extra, a.out, a.timeout)[0]
    return rc


if __name__ == "__main__":
    sys.exit(main())
