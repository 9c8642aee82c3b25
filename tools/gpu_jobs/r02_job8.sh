# round 2, job 8 (1 GPU): full GPU suite with the rebuilt reference (channels as in its pytest.ini), sanitizer on the new kernels,
# bench after the emit selection / forward ring
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu.log; grep -n "^FAILED\|^ERROR" gpurun_out/r02_pytest_gpu.log | head; grep -n "AssertionError" -B3 -A12 gpurun_out/r02_pytest_gpu.log | head -80
cp gpurun_out/refsuite_b200.txt gpurun_out/r02_refsuite_b200.txt 2>/dev/null; cp gpurun_out/refsuite_b200_dropin_stats.json gpurun_out/r02_refsuite_dropin_stats.json 2>/dev/null
head -1 gpurun_out/r02_refsuite_b200.txt; grep "^FAILED" gpurun_out/r02_refsuite_b200.txt | cut -c1-220 | head -20
# compute-sanitizer on a small fwd+bwd (dense + absgrad + backgrounds), memcheck and racecheck
cat > /tmp/san.py <<PY
import torch, numpy as np, sys
sys.path.insert(0, ".")
import gsplat_b200
from tests import scene
sc = scene.make_scene(n_max=6000, sh_degree=3); W, H = 160, 96
Ks = scene.rescale_K(sc["Ks"], sc["width"], sc["height"], W, H)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
P = {k: t(sc[k]).requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "sh")}
P["scales"].data *= 4
for absgrad in (False, True):
    rc, ra, meta = gsplat_b200.rasterization(P["means"], P["quats"], P["scales"], P["opacities"], P["sh"], t(sc["viewmats"][:2]), t(Ks[:2]), W, H, sh_degree=3, packed=False, absgrad=absgrad, backgrounds=torch.rand(2, 3, device="cuda"))
    tgt = torch.rand_like(rc)
    loss = gsplat_b200.l1_loss(rc, tgt) + gsplat_b200.ssim_loss(rc.permute(0, 3, 1, 2), tgt.permute(0, 3, 1, 2)) + ra.mean()
    loss.backward()
torch.cuda.synchronize(); print("sanitizer workload ok", float(loss))
PY
for tool in memcheck racecheck; do compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | tail -4; done > gpurun_out/r02_compute_sanitizer.txt 2>&1; cat gpurun_out/r02_compute_sanitizer.txt
python bench.py --steps 20 --warmup 5 --no-trainer > gpurun_out/r02_v3_bench.json 2> gpurun_out/r02_v3_bench.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r02_v3_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "bwd", round(d["roofline"]["ms"],4), "fwd", round(d["roofline"]["raster_fwd"]["ms"],4), "stock", d["ref_cuda_stock"]["ms_per_step"], "big_s", d["big_s"]["ms_per_step"], d["big_s"]["fwd_only_ms"])
PY
