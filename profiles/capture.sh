#!/bin/bash
# Round-1 evidence capture (run under gpurun on ONE B200): launch lists + full ncu sections of the dominant kernels.
# Every ncu command is preceded by the same command without ncu (must exit 0).
set -u
R=${1:-r01b}
O=gpurun_out
mkdir -p $O
DEC="python profiles/run_phase.py --phase decode --batch 64 --steps 2"
PRE="python profiles/run_phase.py --phase prefill --batch 64"
$DEC > $O/plain_dec.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/${R}_decode_launches_ncu.csv $DEC > $O/ncu_dec.log 2>&1
$PRE > $O/plain_pre.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/${R}_prefill_b64_launches_ncu.csv $PRE > $O/ncu_pre.log 2>&1
# full sections: decode gate||up GEMM (3rd GEMM launch of a layer), decode attention, lm_head; prefill GEMM and tcgen05 attention
$DEC > $O/plain_dec2.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tcgen05_kernel -s 2 -c 2 -o $O/${R}_decode_gemm -f $DEC > $O/ncu_dec_gemm.log 2>&1
$DEC > $O/plain_dec3.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_decode_v3 -s 1 -c 1 -o $O/${R}_decode_attn -f $DEC > $O/ncu_dec_attn.log 2>&1
$PRE > $O/plain_pre2.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tcgen05_kernel -s 120 -c 2 -o $O/${R}_prefill_gemm -f $PRE > $O/ncu_pre_gemm.log 2>&1
$PRE > $O/plain_pre3.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_prefill_tc -s 30 -c 1 -o $O/${R}_prefill_attn -f $PRE > $O/ncu_pre_attn.log 2>&1
ls -la $O/*.ncu-rep
