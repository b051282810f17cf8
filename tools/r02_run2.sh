TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29531 tools/mg_selftest.py > gpurun_out/r02e_selftest.out 2> gpurun_out/r02e_selftest.err; echo selftest rc=$?
grep -v "^\[mg_selftest\] ok" gpurun_out/r02e_selftest.err | grep -v "^$" | tail -8; grep "mg_selftest" gpurun_out/r02e_selftest.err | tail -2; tail -2 gpurun_out/r02e_selftest.out
GFFM_TRACE=1 GFFM_TRACE_LAST=110 timeout 200 $TR --master-port 29541 bench.py --gpus 2 --steps 12 --warmup 3 --transport p2p_push --no-e2e --no-parity --no-cpu > gpurun_out/r02e_trace_push.json 2> gpurun_out/r02e_trace_push.err; echo trace rc=$?
for t in p2p_push p2p_planes; do timeout 200 $TR --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --transport $t --no-cpu > gpurun_out/r02e_bench2_$t.json 2> gpurun_out/r02e_bench2_$t.err; echo $t rc=$?; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02e_bench2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['ms_per_step'],3), 'gemm/step', r['gemm_ms_per_step_per_rank'], 'shards', d['config']['shards_match_local_product_on_all_ranks'], 'parity', d['parity_check'] and d['parity_check']['match'], d['clocks']['sm_mhz'], d['config']['multi_gpu_info'], 'e2e', d['e2e'] and d['e2e'].get('ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
PY
