N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NCCL_DEBUG=INFO timeout 300 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02f_bench${N}_tune.json 2> gpurun_out/r02f_bench${N}_tune.err; echo tune rc=$?
for t in p2p_planes p2p_push; do GFFM_TRACE=1 GFFM_TRACE_LAST=150 timeout 200 $TR --master-port 29542 bench.py --gpus $N --steps 12 --warmup 3 --transport $t --no-e2e --no-parity --no-cpu > gpurun_out/r02f_trace${N}_$t.json 2> gpurun_out/r02f_trace${N}_$t.err; echo $t rc=$?; done
rm -f gpurun_out/nccl_debug.*.log.keep; ls gpurun_out/nccl_debug.* 2>/dev/null | head -3
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02f_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['ms_per_step'],3), 'gemm/step', r['gemm_ms_per_step_per_rank'], 'shards', d['config']['shards_match_local_product_on_all_ranks'], 'parity', d['parity_check'] and d['parity_check']['match'], d['clocks']['sm_mhz'], d['config']['multi_gpu_transport'], d['config']['warmup_trials_ms'], 'e2e', d['e2e'] and d['e2e'].get('ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
PY
