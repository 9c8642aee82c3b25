"""pytest plugin (``-p gsb200_refsuite_plugin``) for running the REFERENCE's own test-suite
(baseline/_ref/reference_suite, copied verbatim by baseline/install_ref.py) against a chosen backend:

  GSB200_DROPIN=1   gsplat_b200.dropin.apply() is called before the reference's test modules are imported, so
                    every ``gsplat.rasterization`` / ``gsplat.cuda._wrapper.<op>`` the tests resolve is ours
                    (INTEGRATION.md section A); the Python twins they compare against stay the reference's.
  GSB200_DROPIN=0   the unmodified reference (harness sanity: the same selection must pass there too).

Test infrastructure only.
"""
import os
import sys


def pytest_configure(config):
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(os.path.dirname(here))
    if root not in sys.path:
        sys.path.append(root)  # behind the suite's own rootdir: ``tests`` must stay the reference's package
    from oracle import refcuda

    refcuda.import_package()
    if os.environ.get("GSB200_DROPIN", "0") == "1":
        from gsplat_b200 import dropin

        info = dropin.apply(losses=True)
        print(f"[gsb200] drop-in applied: {info}")


def pytest_sessionfinish(session, exitstatus):
    """How many calls of every rebound name ran on the b200 kernels / were handed to the stock implementation."""
    if os.environ.get("GSB200_DROPIN", "0") != "1":
        return
    import json

    from gsplat_b200 import dropin

    out = os.environ.get("GSB200_DROPIN_STATS")
    if out:
        with open(out, "w") as f:
            json.dump(dropin.stats, f, indent=1, sort_keys=True)
    print("[gsb200] drop-in call statistics:", {k: (v["b200"], v["stock"]) for k, v in dropin.stats.items()})
