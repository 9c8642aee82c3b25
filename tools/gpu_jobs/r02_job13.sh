# round 2, job 13 (1 GPU): one-call capacity-sized intersection stage with the totals read overlapped -- parity, A/B, launch list
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference_cuda.py -m gpu -q -x --tb=short 2>&1 | tail -5
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 30 --warmup 5 --no-trainer --no-cpu-baseline > gpurun_out/r02_v6_bench_$name.json 2> gpurun_out/r02_v6_bench_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_v6_bench_$name.json").read().strip().splitlines()[-1])
print("$name", "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "bwd", round(d["roofline"]["ms"],4), "fwd", round(d["roofline"]["raster_fwd"]["ms"],4), "stock", round(d["ref_cuda_stock"]["ms_per_step"],3), "big_s", round(d["big_s"]["ms_per_step"],3), round(d["big_s"]["fwd_only_ms"],3))
PY
}
run spec GSB200_X=0
run nospec GSB200_ISECT_SPECULATE=0
run spec_norows GSB200_ROW_RECORDS=0
cat > /tmp/san.py <<PY
import torch, numpy as np, sys
sys.path.insert(0, ".")
import gsplat_b200
from tests import scene
sc = scene.make_scene(n_max=6000, sh_degree=3); W, H = 160, 96
Ks = scene.rescale_K(sc["Ks"], sc["width"], sc["height"], W, H)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
P = {k: t(sc[k]).requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "sh")}
for it, mul in enumerate((1.0, 1.0, 4.0, 0.5)):
    P["scales"].data *= mul
    rc, ra, meta = gsplat_b200.rasterization(P["means"], P["quats"], P["scales"], P["opacities"], P["sh"], t(sc["viewmats"][:2]), t(Ks[:2]), W, H, sh_degree=3, packed=False, absgrad=(it % 2 == 1), backgrounds=torch.rand(2, 3, device="cuda"))
    tgt = torch.rand_like(rc)
    loss = gsplat_b200.l1_loss(rc, tgt) + gsplat_b200.ssim_loss(rc.permute(0, 3, 1, 2), tgt.permute(0, 3, 1, 2)) + ra.mean()
    loss.backward()
    _ = meta["isect_ids"]
torch.cuda.synchronize(); print("sanitizer workload ok", float(loss.detach()))
PY
compute-sanitizer --tool memcheck --print-limit 10 python /tmp/san.py > gpurun_out/r02_sanitizer_memcheck3.txt 2>&1; tail -2 gpurun_out/r02_sanitizer_memcheck3.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v6_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-trainer > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/r02_v6_launches.csv
