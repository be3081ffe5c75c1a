#!/bin/bash
# The C driver with --profile on a slice of a workload under N ranks: tools/driver_profile_n.sh <workload> <queries> <steps> <ranks> [flags]
W=$1; Q=$2; STEPS=$3; N=$4; shift 4
IN=/tmp/mcq_in_${W}_${Q}
python - <<PY
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from chain_files import make_chain_inputs
from mcqueen_b200 import synth
paths, L, _ = make_chain_inputs("$IN", "$W", seed=20261019, query_model="dsim", Q=$Q)
open("$IN/L", "w").write(str(L))
PY
L=$(cat $IN/L)
D=mcqueen_b200/host/mcq_dmodel
A="-H $IN/query.HLA.csv -l $L -n $STEPS -s 1000000 $IN/ref.fasta $IN/query.fasta"
if [ "$N" = "1" ]; then
  $D --seed 1 --gsl_seed 1 --profile "$@" $A /tmp/mcq_out_1 2>&1 | grep -E "^profile|MCMC loop|final log"
else
  python -m torch.distributed.run --no-python --nnodes=1 --master-addr 127.0.0.1 --master-port 29533 --nproc-per-node $N \
    $D --seed 1 --gsl_seed 1 --profile "$@" $A /tmp/mcq_out_$N 2>&1 | grep -E "^profile|MCMC loop|final log" | sort | uniq -c | sort -k2 | head -40
fi
