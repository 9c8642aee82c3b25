# round 2, job 11 (1 GPU): narrow-key intersection pipeline -- parity, A/B against the wide keys, cost of the host read,
# racecheck of the CTA-barrier variants, launch list of the new step
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference_cuda.py -m gpu -q -x --tb=short 2>&1 | tail -5
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 30 --warmup 5 --no-trainer --no-cpu-baseline > gpurun_out/r02_v4_bench_$name.json 2> gpurun_out/r02_v4_bench_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_v4_bench_$name.json").read().strip().splitlines()[-1])
print("$name", "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "bwd", round(d["roofline"]["ms"],4), "fwd", round(d["roofline"]["raster_fwd"]["ms"],4), "stock", round(d["ref_cuda_stock"]["ms_per_step"],3), "big_s", round(d["big_s"]["ms_per_step"],3), round(d["big_s"]["fwd_only_ms"],3))
PY
}
run narrow GSB200_X=0
run wide GSB200_ISECT_WIDE=1
run narrow_nosync GSB200_MEASURE_NO_SYNC=1
run narrow_f32target GSB200_E2E_TARGET=f32
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
GSB200_FWD_PIPE=0 GSB200_BWD_PIPE=0 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 10 python /tmp/san.py > gpurun_out/r02_sanitizer_racecheck_barrier.txt 2>&1; tail -2 gpurun_out/r02_sanitizer_racecheck_barrier.txt
compute-sanitizer --tool memcheck --print-limit 10 python /tmp/san.py > gpurun_out/r02_sanitizer_memcheck2.txt 2>&1; tail -1 gpurun_out/r02_sanitizer_memcheck2.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v4_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-trainer > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/r02_v4_launches.csv
