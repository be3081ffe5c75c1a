#!/bin/bash
# One 8-GPU session: C4 sharded over 8 and 4 GPUs, C5 over 8, and the C driver's chain at C4 on 8 GPUs.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/bench_c4_r1_n8.json 2> gpurun_out/bench_c4_r1_n8.err; echo "c4 n8 rc=$?"
$TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 50 --warmup 3 > gpurun_out/bench_c4_r1_n4.json 2> gpurun_out/bench_c4_r1_n4.err; echo "c4 n4 rc=$?"
$TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --workload C5 --steps 10 --warmup 3 --no-extras > gpurun_out/bench_c5_r1_n8.json 2> gpurun_out/bench_c5_r1_n8.err; echo "c5 n8 rc=$?"
python - <<PY
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from chain_files import make_chain_inputs
make_chain_inputs("/tmp/c4in", "C4", seed=20261019)
PY
D=mcqueen_b200/host/mcq_dmodel
A="-H /tmp/c4in/query.HLA.csv -l 500 -n 1000 -s 100000 /tmp/c4in/ref.fasta /tmp/c4in/query.fasta"
for extra in "" "--dense_esc"; do
  echo "== 8 GPUs $extra"
  $TR --no-python --nproc-per-node 8 --master-port 29524 $D --seed 1 --gsl_seed 1 $extra $A /tmp/c4out8 2>&1 | grep -E "MCMC loop|final log" | sort | uniq -c
done
