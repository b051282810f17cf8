TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
MG_SELFTEST_TRANSPORTS=p2p_push MG_SELFTEST_TIMING=0 timeout 200 $TR --master-port 29531 tools/mg_selftest.py > gpurun_out/r02j_selftest_a.out 2> gpurun_out/r02j_selftest_a.err; echo selftest push rc=$?
grep "FAIL" gpurun_out/r02j_selftest_a.err | tail -5; tail -1 gpurun_out/r02j_selftest_a.out
GFFM_MG_PUSH_CE_PEERS=1 MG_SELFTEST_TRANSPORTS=p2p_push MG_SELFTEST_TIMING=0 timeout 200 $TR --master-port 29532 tools/mg_selftest.py > gpurun_out/r02j_selftest_b.out 2> gpurun_out/r02j_selftest_b.err; echo selftest push+ce rc=$?
grep "FAIL" gpurun_out/r02j_selftest_b.err | tail -5; tail -1 gpurun_out/r02j_selftest_b.out
timeout 200 $TR --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --transport p2p_push --no-cpu --no-e2e > gpurun_out/r02j_bench2_push.json 2> gpurun_out/r02j_bench2_push.err; echo rc=$?
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02j_bench2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['ms_per_step'],3), 'gemm/step', r['gemm_ms_per_step_per_rank'], 'shards', d['config']['shards_match_local_product_on_all_ranks'], 'parity', d['parity_check'] and d['parity_check']['match'], d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
