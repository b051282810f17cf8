N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NCCL_DEBUG=INFO timeout 300 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --transport p2p_push,nccl_bcast > gpurun_out/r02i_bench${N}_tune.json 2> gpurun_out/r02i_bench${N}_tune.err; echo tune rc=$?
GFFM_TRACE=1 GFFM_TRACE_LAST=150 timeout 200 $TR --master-port 29542 bench.py --gpus $N --steps 12 --warmup 3 --transport p2p_push --no-e2e --no-parity --no-cpu > gpurun_out/r02i_trace${N}_p2p_push.json 2> gpurun_out/r02i_trace${N}_p2p_push.err; echo trace rc=$?
GFFM_PUSH_CTAS_PER_SM=2 timeout 200 $TR --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 --transport p2p_push --no-e2e --no-parity --no-cpu > gpurun_out/r02i_bench${N}_push2.json 2> gpurun_out/r02i_bench${N}_push2.err; echo push2 rc=$?
GFFM_PUSH_CTAS_PER_SM=8 timeout 200 $TR --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 --transport p2p_push --no-e2e --no-parity --no-cpu > gpurun_out/r02i_bench${N}_push8.json 2> gpurun_out/r02i_bench${N}_push8.err; echo push8 rc=$?
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02i_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['ms_per_step'],3), 'gemm/step', r['gemm_ms_per_step_per_rank'], 'shards', d['config']['shards_match_local_product_on_all_ranks'], 'parity', d['parity_check'] and d['parity_check']['match'], d['clocks']['sm_mhz'], d['config']['multi_gpu_transport'], d['config']['warmup_trials_ms'], 'e2e', d['e2e'] and d['e2e'].get('ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
PY
grep "trace r3" gpurun_out/r02i_trace${N}_p2p_push.err | grep -v bcast | grep "splitB\|wait:free\|wait:staged\|gemm  *[03]" | tail -16
