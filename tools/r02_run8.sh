N=${1:-8}
TAG=${TAG:-r02o}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NCCL_DEBUG=INFO timeout 300 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench${N}_tune.json 2> gpurun_out/${TAG}_bench${N}_tune.err; echo tune rc=$?
timeout 200 $TR --master-port 29543 bench.py --gpus $N --steps 40 --warmup 5 --transport p2p_push --no-e2e --no-parity --no-cpu > gpurun_out/${TAG}_bench${N}_push40.json 2> gpurun_out/${TAG}_bench${N}_push40.err; echo push40 rc=$?
GFFM_TRACE=1 GFFM_TRACE_LAST=150 timeout 200 $TR --master-port 29542 bench.py --gpus $N --steps 12 --warmup 3 --transport p2p_push --no-e2e --no-parity --no-cpu > gpurun_out/${TAG}_trace${N}.json 2> gpurun_out/${TAG}_trace${N}.err; echo trace rc=$?
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/%s_*.json' % os.environ.get('TAG','r02o'))):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['ms_per_step'],3), 'gemm/step', r['gemm_ms_per_step_per_rank'], 'shards', d['config']['shards_match_local_product_on_all_ranks'], 'parity', d['parity_check'] and d['parity_check']['match'], d['clocks']['sm_mhz'], d['config']['multi_gpu_transport'], d['config']['warmup_trials_ms'], 'e2e', d['e2e'] and d['e2e'].get('ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
PY
for r in 0 3; do grep "trace r$r\]" gpurun_out/${TAG}_trace${N}.err | head -170 | grep "gemm  *3 \|splitB\|wait:staged\|wait:free\|push " | cut -c16-90 | tail -24; done
