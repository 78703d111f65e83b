for c in read copy; do  # NPROC from env
  PD_PEER_MODE=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NPROC:-2} --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py > gpurun_out/sharded${NPROC:-2}_peer_$c.log 2>&1; echo rc=$?
  grep ms_peer_hpsi gpurun_out/sharded${NPROC:-2}_peer_$c.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$c', d['ms_peer_hpsi'], d['ms_peer_hpsi_staged'], d['ms_local_hpsi'], d['peer_vs_exchange_maxabs'])"
  grep ms_per_hpsi gpurun_out/sharded${NPROC:-2}_peer_$c.log
done
