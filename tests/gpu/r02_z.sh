#!/bin/bash
# round 2, third session: the FINAL record sweep (x-limited rings in the search, device finalresults/contributions) -- GPU tests, smoke, default bench (C4 + C2 secondary),
# reference arm, per-config bench lines, ncu launch lists of C4/C2/C3/C5, `ncu --set full` of the dominant kernels
# (k_tile_search on C4 and C2, k_pairs / k_pair_random on C3).  Steps are skipped once the time budget is spent.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s); LIMIT=${1:-1200}
left() { [ $(( $(date +%s) - T0 )) -lt $LIMIT ]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02z_bench_default.json 2> gpurun_out/r02z_bench_default.err; tail -3 gpurun_out/r02z_bench_default.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02z_bench_default.json").read().strip().splitlines()[-1])
    print("workload", d["config"]["workload"][:40], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "job", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["job"].items() if k!="what"})
    print("launches/frame", d["launches_per_frame"], "host submit/wait/cpu ms", round(d["host_submit_ms_per_step"],2), round(d["host_wait_ms_per_step"],2), round(d["host_cpu_ms_per_step"],2), "of", round(d["ms_per_step"],2), "clocks", d["clocks"])
    print("cpu", d["cpu_baseline"])
    print("roofline", {k:v for k,v in d["roofline"].items() if k in ("kernel","achieved","frac","kernel_ms_per_frame","kernel_share_of_frame","pair_evals_per_frame","frames_per_launch","alu_view")})
    print("hbm kernel", d["roofline_hbm_kernel"])
    s=d["secondary"]; print("secondary C2 value", round(s["value"],1), "e2e", round(s["e2e"]["value"],1), "launches/frame", s["launches_per_frame"], "submit/wait", round(s["host_submit_ms_per_step"],2), round(s["host_wait_ms_per_step"],2), "of", round(s["ms_per_step"],2))
except Exception as e:
    print("default bench failed", e)
PY
left && { timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02z_bench_reference_arm.json 2> gpurun_out/r02z_bench_reference_arm.err; cut -c1-300 gpurun_out/r02z_bench_reference_arm.json; echo; }
# ---- ncu launch lists: one stream (= one batch in flight) so that the list is the serial order of a batch
LL="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-job --no-secondary --no-hbm-kernel --streams 1"
for cfg in C4 C2 C3; do
  left && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02z_launches_$cfg.csv \
     python bench.py --config $cfg $LL --frames-per-step 16 > gpurun_out/r02z_ncu_ll_$cfg.log 2>&1
done
left && timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02z_launches_C5.csv \
     python bench.py --config C5 $LL --frames-per-step 4 > gpurun_out/r02z_ncu_ll_C5.log 2>&1
# ---- full captures of the dominant kernels
left && { timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_tile_search -s 10 -c 2 -f -o gpurun_out/r02z_prof_search_C4 \
   python bench.py --config C4 $LL --frames-per-step 16 > gpurun_out/r02z_ncu_full_C4.log 2>&1; tail -1 gpurun_out/r02z_ncu_full_C4.log | cut -c1-120; }
left && { timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_tile_search -s 10 -c 2 -f -o gpurun_out/r02z_prof_search_C2 \
   python bench.py --config C2 $LL --frames-per-step 16 > gpurun_out/r02z_ncu_full_C2.log 2>&1; tail -1 gpurun_out/r02z_ncu_full_C2.log | cut -c1-120; }
# ---- per-config bench lines (no secondary / job: those belong to the default line)
run() { C=$1; shift; left || { echo "skip $C (time)"; return; }; timeout 500 python bench.py --config $C --no-secondary --no-hbm-kernel "$@" > gpurun_out/r02z_bench_$C.json 2> gpurun_out/r02z_bench_$C.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02z_bench_$C.json").read().strip().splitlines()[-1])
    cb=d.get("cpu_baseline") or {}
    print("$C value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "cpu", round(cb.get("value",0),3), "equal", cb.get("counts_equal_device"), "| kernel", d["roofline"]["kernel"], "ms/frame", round(d["roofline"]["kernel_ms_per_frame"],4), "share", round(d["roofline"]["kernel_share_of_frame"],3), "pair-evals/s", "%.3g"%d["roofline"]["pair_evals_per_s"], "launches/frame", round(d["launches_per_frame"],2), "submit/wait", round(d["host_submit_ms_per_step"],2), round(d["host_wait_ms_per_step"],2), "of", round(d["ms_per_step"],2))
except Exception as e:
    print("$C failed", e); print(open("gpurun_out/r02z_bench_$C.err").read()[-600:])
PY
}
run C3 --steps 5 --cpu-frames 4
run C2urea --steps 8 --cpu-frames 16
run C5 --steps 3 --no-cpu-baseline --no-job
echo "elapsed $(( $(date +%s) - T0 )) s"
