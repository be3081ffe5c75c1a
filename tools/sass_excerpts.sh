#!/bin/bash
# Writes profiles/sass_r2.txt: per kernel the registers / shared memory (cuobjdump -res-usage) and the instruction mix of
# its SASS (opcode histogram), and the body of one site of the stored forward sweep.  Runs anywhere nvcc's tools are
# (no GPU needed): the listing is of the library as built by mcqueen_b200/build.py.
LIB=mcqueen_b200/libmcq_b200.so
OUT=profiles/sass_r2.txt
KERNELS="_Z9k_forwardILi8ELb1ELb0ELb1ELb0EEv7FwdArgs _Z9k_forwardILi8ELb1ELb0ELb0ELb0EEv7FwdArgs _Z10k_backwardILi8ELb1ELb0EEv7BwdArgs _Z12k_forward_ckILi8ELb1EEv7FwdArgs _Z13k_backward_ckILi8ELb1EEv7BwdArgs _Z8k_columnILi8ELb1EEv7FwdArgs _Z5k_dot7DotArgs _Z12k_commit_all10CommitArgs _Z15k_emission_both8EmisArgsj _Z14k_sum_exchangePKdiPd7PeerBoxyPiP11McqHostSloty _Z8k_gatherPKaiS0_PKiS0_PKhPhiiiiPi _Z9k_hammingILi32ELi2EEvPKymiS1_iiPtm _Z12k_select_rowPKtmiiiii9SelectOut _Z11k_consensusILi11EEvPKyiPKiiiiPa"
{
echo "SASS of $LIB (sm_100a), $(nvcc --version | tail -1)"
echo "flags: -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17"
echo
for k in $KERNELS; do
  echo "== $(echo $k | c++filt)"
  cuobjdump -res-usage -fun $k $LIB 2>/dev/null | grep -E "REG:" | head -1
  cuobjdump -sass -fun $k $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | awk '{ op=$2; if (op ~ /^@/) op=$3; sub(/;$/,"",op); n[op]++ } END { for (o in n) printf "%6d %s\n", n[o], o }' | sort -rn | head -28 | paste -sd' ' | fold -w 150
  echo
done
echo "== one site of the stored forward sweep (k_forward<8, FULL, !PRE, STORE>): from the wait for the staged rows to the column store"
cuobjdump -sass -fun _Z9k_forwardILi8ELb1ELb0ELb1ELb0EEv7FwdArgs $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\*[0-9a-f]*\*/##g; s/ *$//' | awk '/DEPBAR.LE SB0, 0x2/{c++} c==2' | awk '/DEPBAR.LE SB0, 0x2/{c++} c<2' | cut -c20-110
} > $OUT
wc -l $OUT
