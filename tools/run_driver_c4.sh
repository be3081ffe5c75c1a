#!/bin/bash
# Tuning aid: the C driver at C4 scale on 1 GPU and on N GPUs (torchrun --no-python).
set -e
N=${1:-2}
STEPS=${2:-300}
python - <<PY
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from chain_files import make_chain_inputs
paths, L, _ = make_chain_inputs("/tmp/c4in", "C4", seed=20261019)
print(paths, L)
PY
D=mcqueen_b200/host/mcq_dmodel
A="-H /tmp/c4in/query.HLA.csv -l 500 -n $STEPS -s 100000 /tmp/c4in/ref.fasta /tmp/c4in/query.fasta"
echo "== 1 GPU"; SECONDS=0; $D --seed 1 --gsl_seed 1 $A /tmp/c4out1 | grep -E "Total log|MCMC loop|final log|accept"; echo "wall ${SECONDS}s"
echo "== $N GPUs (peer exchange)"; python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 $D --seed 1 --gsl_seed 1 $A /tmp/c4outN 2>&1 | grep -E "Total log|MCMC loop|final log" | sort | uniq -c
echo "== 1 GPU, --dense_esc"; $D --seed 1 --gsl_seed 1 --dense_esc $A /tmp/c4out1s | grep -E "Total log|MCMC loop|final log"
echo "== 1 GPU, --eager_refill"; $D --seed 1 --gsl_seed 1 --dense_esc --eager_refill $A /tmp/c4out1e | grep -E "Total log|MCMC loop|final log"
