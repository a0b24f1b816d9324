#!/bin/bash
# final check of the tree as committed: GPU suite, smoke, default bench line, launch list + full capture of the final search kernel (C4)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02fin_bench_default.json 2> gpurun_out/r02fin_bench_default.err; tail -3 gpurun_out/r02fin_bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02fin_bench_default.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "clocks", d["clocks"], "job", round(d["job"]["frames_per_s"],1), round(d["job"]["wall_s"],3), "cpu", round(d["cpu_baseline"]["value"],3), d["cpu_baseline"]["counts_equal_device"])
r=d["roofline"]; print("roofline", r["kernel"], round(r["kernel_ms_per_frame"],4), round(r["kernel_share_of_frame"],3), "frac", round(r["frac"],3), "pe/frame %.4g"%r["pair_evals_per_frame"], "alu %.4g"%r["alu_view"]["kernel_pair_evals_per_s"])
s=d["secondary"]; print("secondary C2", round(s["value"],1), round(s["e2e"]["value"],1))
PY
LL="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-job --no-secondary --no-hbm-kernel --streams 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02fin_launches_C4.csv python bench.py --config C4 $LL --frames-per-step 16 > gpurun_out/r02fin_ncu_ll_C4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile_search -s 10 -c 2 -f -o gpurun_out/r02fin_prof_search_C4 python bench.py --config C4 $LL --frames-per-step 16 > gpurun_out/r02fin_ncu_full_C4.log 2>&1; tail -1 gpurun_out/r02fin_ncu_full_C4.log | cut -c1-100
timeout 120 python bench.py --config C5 --steps 3 --no-cpu-baseline --no-job --no-secondary --no-hbm-kernel > gpurun_out/r02fin_bench_C5.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02fin_bench_C5.json').read().strip().splitlines()[-1]); print('C5 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
timeout 60 python bench.py --config C2 --steps 8 --cpu-frames 16 --no-job --no-secondary --no-hbm-kernel > gpurun_out/r02fin_bench_C2.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02fin_bench_C2.json').read().strip().splitlines()[-1]); print('C2 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['counts_equal_device'])"
