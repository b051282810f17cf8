TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 tools/mg_selftest.py > gpurun_out/r02h_selftest.out 2> gpurun_out/r02h_selftest.err; echo selftest rc=$?
grep "FAIL" gpurun_out/r02h_selftest.err | tail -8; grep -c "ok  " gpurun_out/r02h_selftest.err; tail -1 gpurun_out/r02h_selftest.out
for t in p2p_push p2p_planes; do timeout 200 $TR --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --transport $t --no-cpu --no-e2e > gpurun_out/r02h_bench2_$t.json 2> gpurun_out/r02h_bench2_$t.err; echo $t rc=$?; done
GFFM_TRACE=1 GFFM_TRACE_LAST=60 timeout 200 $TR --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --transport p2p_push --no-e2e --no-parity --no-cpu > gpurun_out/r02h_trace_push.json 2> gpurun_out/r02h_trace_push.err; echo trace rc=$?
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02h_bench2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['ms_per_step'],3), 'gemm/step', r['gemm_ms_per_step_per_rank'], 'shards', d['config']['shards_match_local_product_on_all_ranks'], 'parity', d['parity_check'] and d['parity_check']['match'], d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
grep "trace r0" gpurun_out/r02h_trace_push.err | grep -v bcast | tail -24
