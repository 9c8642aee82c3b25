# round 2, job 10 (1 GPU): parity subset after the racecheck hygiene changes + racecheck with the full hazard list
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference_cuda.py -m gpu -q -x --tb=short 2>&1 | tail -5
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
torch.cuda.synchronize(); print("sanitizer workload ok", float(loss.detach()))
PY
compute-sanitizer --tool memcheck --print-limit 10 python /tmp/san.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; tail -3 gpurun_out/r02_sanitizer_memcheck.txt
compute-sanitizer --tool racecheck --racecheck-report all --print-limit 40 python /tmp/san.py > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; tail -3 gpurun_out/r02_sanitizer_racecheck.txt
grep -n "hazard detected\|Write Thread\|Read Thread\|in function\|raster\|ssim\|isect" gpurun_out/r02_sanitizer_racecheck.txt | awk '{$1=$1};1' | sort | uniq -c | sort -rn | head -40
compute-sanitizer --tool synccheck --print-limit 10 python /tmp/san.py > gpurun_out/r02_sanitizer_synccheck.txt 2>&1; tail -2 gpurun_out/r02_sanitizer_synccheck.txt
