TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 700 $TR --master-port 29531 tools/mg_selftest.py > gpurun_out/r02d_selftest.out 2> gpurun_out/r02d_selftest.err; echo selftest rc=$?
grep -v "^\[mg_selftest\] ok" gpurun_out/r02d_selftest.err | grep -v "^$" | tail -12; tail -2 gpurun_out/r02d_selftest.out
GFFM_TRACE=1 GFFM_TRACE_LAST=140 timeout 300 $TR --master-port 29541 bench.py --gpus 2 --steps 6 --warmup 3 --transport p2p_planes --no-e2e --no-parity --no-cpu > gpurun_out/r02d_trace_p2p.json 2> gpurun_out/r02d_trace_p2p.err; echo trace rc=$?
for t in p2p_planes nccl_planes nccl_bcast; do timeout 300 $TR --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --transport $t --no-cpu > gpurun_out/r02d_bench2_$t.json 2> gpurun_out/r02d_bench2_$t.err; echo $t rc=$?; done
GFFM_MG_FLAGS=kernel timeout 300 $TR --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 --transport p2p_planes --no-e2e --no-cpu > gpurun_out/r02d_bench2_p2p_kernelflags.json 2> gpurun_out/r02d_bench2_p2p_kernelflags.err; echo kernelflags rc=$?
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02d_bench2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['ms_per_step'],3), 'gemm/step', r['gemm_ms_per_step_per_rank'], 'shards', d['config']['shards_match_local_product_on_all_ranks'], 'parity', d['parity_check'] and d['parity_check']['match'], d['clocks']['sm_mhz'], d['config']['multi_gpu_info'], 'e2e', d['e2e'] and d['e2e'].get('ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
PY
