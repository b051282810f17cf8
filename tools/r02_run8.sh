N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { tag=$1; shift; env "$@" timeout 200 $TR --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 --transport p2p_push --no-e2e --no-parity --no-cpu > gpurun_out/r02k_bench${N}_$tag.json 2> gpurun_out/r02k_bench${N}_$tag.err; echo $tag rc=$?; }
run sms20_ce0 GFFM_MG_PUSH_SMS=20 GFFM_MG_PUSH_CE_PEERS=0
run sms20_ce3 GFFM_MG_PUSH_SMS=20 GFFM_MG_PUSH_CE_PEERS=3
run sms12_ce3 GFFM_MG_PUSH_SMS=12 GFFM_MG_PUSH_CE_PEERS=3
run sms12_ce5 GFFM_MG_PUSH_SMS=12 GFFM_MG_PUSH_CE_PEERS=5
GFFM_TRACE=1 GFFM_TRACE_LAST=150 GFFM_MG_PUSH_SMS=20 GFFM_MG_PUSH_CE_PEERS=3 timeout 200 $TR --master-port 29542 bench.py --gpus $N --steps 12 --warmup 3 --transport p2p_push --no-e2e --no-parity --no-cpu > gpurun_out/r02k_trace${N}.json 2> gpurun_out/r02k_trace${N}.err; echo trace rc=$?
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02k_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['ms_per_step'],3), 'gemm/step', r['gemm_ms_per_step_per_rank'], 'shards', d['config']['shards_match_local_product_on_all_ranks'], d['clocks']['sm_mhz'], d['config']['multi_gpu_transport'])
    except Exception as e: print(f, 'ERR', e)
PY
grep "trace r3" gpurun_out/r02k_trace${N}.err | grep -v bcast | grep "splitB\|wait:free\|cepush\|gemm  *[03]" | tail -14
grep "trace r0" gpurun_out/r02k_trace${N}.err | grep -v bcast | grep "splitB\|wait:free\|push " | tail -12
