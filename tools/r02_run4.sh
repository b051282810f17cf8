TAG=${TAG:-r02r}
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
MG_SELFTEST_TRANSPORTS=p2p_raw MG_SELFTEST_TIMING=0 timeout 40 $TR --master-port 29531 tools/mg_selftest.py > gpurun_out/${TAG}_selftest_a.out 2> gpurun_out/${TAG}_selftest_a.err; echo selftest raw rc=$?
grep "FAIL" gpurun_out/${TAG}_selftest_a.err | tail -5; tail -1 gpurun_out/${TAG}_selftest_a.out
GFFM_MG_ROOT_FREE_MIN=4 MG_SELFTEST_TRANSPORTS=p2p_raw MG_SELFTEST_TIMING=0 timeout 40 $TR --master-port 29532 tools/mg_selftest.py > gpurun_out/${TAG}_selftest_b.out 2> gpurun_out/${TAG}_selftest_b.err; echo selftest raw rootfree rc=$?
grep "FAIL" gpurun_out/${TAG}_selftest_b.err | tail -5; tail -1 gpurun_out/${TAG}_selftest_b.out
timeout 50 $TR --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_bench${N}_tune.json 2> gpurun_out/${TAG}_bench${N}_tune.err; echo bench rc=$?
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/%s_bench*.json' % os.environ.get('TAG','r02r'))):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['ms_per_step'],3), 'gemm/step', r['gemm_ms_per_step_per_rank'], 'shards', d['config']['shards_match_local_product_on_all_ranks'], 'parity', d['parity_check'] and d['parity_check']['match'], d['clocks']['sm_mhz'], d['config']['multi_gpu_transport'], d['config']['warmup_trials_ms'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/${TAG}_bench${N}_tune.err
