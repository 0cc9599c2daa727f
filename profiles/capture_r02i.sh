#!/bin/bash
# Round-2 evidence capture (run under gpurun on ONE B200): launch lists + full ncu sections of the kernels the decode step and the
# prefill launch NOW.  Every ncu command is preceded by the same command without ncu (must exit 0).  Outputs: gpurun_out/r02_*.
set -u
R=${1:-r02i}
O=gpurun_out
mkdir -p $O
NCU_T="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
NCU_F="ncu --set full --clock-control none --import-source on --profile-from-start off"
DEC="python profiles/run_phase.py --phase decode --batch 64 --steps 2"
BENCH="python bench.py --steps 1 --warmup 1 --no-configs --no-serving --no-cpu-baseline"

# 1. launch list of the bench command itself: ~5 decode steps of the timed generate() (graph kernel nodes)
$BENCH > $O/${R}_plain_bench.log 2>&1 && \
$NCU_T -s 18000 -c 700 --log-file $O/${R}_bench_launches_ncu.csv $BENCH > $O/${R}_ncu_bench.log 2>&1

# 2. eager decode steps, kernel by kernel (profiler started after the prefill)
$DEC > $O/${R}_plain_dec.log 2>&1 && \
$NCU_T --profile-from-start off --log-file $O/${R}_decode_launches_ncu.csv $DEC > $O/${R}_ncu_dec.log 2>&1
# full sections: the four GEMMs of a layer (qkv, o, gate||up, down: first 7 GEMM launches), the lm_head (73rd), decode attention,
# the statistics sampler, the decode RMSNorm
$NCU_F -k regex:gemm_tcgen05_kernel -c 7 -o $O/${R}_decode_gemm -f $DEC > $O/${R}_ncu_dec_gemm.log 2>&1
$NCU_F -k regex:gemm_tcgen05_kernel -s 72 -c 1 -o $O/${R}_decode_head -f $DEC > $O/${R}_ncu_dec_head.log 2>&1
$NCU_F -k regex:attn_decode_v3 -s 1 -c 1 -o $O/${R}_decode_attn -f $DEC > $O/${R}_ncu_dec_attn.log 2>&1
$NCU_F -k regex:sample_top_p_stats -c 1 -o $O/${R}_decode_sampler -f $DEC > $O/${R}_ncu_dec_sampler.log 2>&1
$NCU_F -k regex:rmsnorm_kernel -s 1 -c 1 -o $O/${R}_decode_rmsnorm -f $DEC > $O/${R}_ncu_dec_rmsnorm.log 2>&1

# 3. prefill launch lists with the current kernels: 224 px x 64, 448 px x 32, 896 px x 8
for cfg in "224 64" "448 32" "896 8"; do
  set -- $cfg
  PRE="python profiles/run_phase.py --phase prefill --image-size $1 --batch $2"
  $PRE > $O/${R}_plain_pre_$1.log 2>&1 && \
  $NCU_T --profile-from-start off --log-file $O/${R}_prefill_$1_b$2_launches_ncu.csv $PRE > $O/${R}_ncu_pre_$1.log 2>&1
done
# full sections at 896 px (where attention matters): SigLIP attention (dh 72, 4096 keys), Gemma attention (dh 256, 4100 keys), main GEMM
PRE="python profiles/run_phase.py --phase prefill --image-size 896 --batch 8"
$NCU_F -k regex:attn_prefill_tc -s 3 -c 1 -o $O/${R}_prefill896_attn72 -f $PRE > $O/${R}_ncu_pre_attn72.log 2>&1
$NCU_F -k regex:attn_prefill_tc -s 30 -c 1 -o $O/${R}_prefill896_attn256 -f $PRE > $O/${R}_ncu_pre_attn256.log 2>&1
$NCU_F -k regex:gemm_pair_kernel -s 84 -c 2 -o $O/${R}_prefill896_gemm -f $PRE > $O/${R}_ncu_pre_gemm.log 2>&1

# 4. summaries (small, tracked copies go to profiles/)
for f in $O/${R}_*_launches_ncu.csv; do python profiles/summarize_launches.py $f > ${f%_ncu.csv}_summary.txt 2>&1; done
for f in $O/${R}_*.ncu-rep; do python profiles/summarize_ncu_rep.py $f > ${f%.ncu-rep}_ncu_full.csv 2>&1; done
ls -la $O/${R}_* | head -60
# keep the two reports whose source pages are quoted in DESIGN.md; the rest is summarised above (gpurun_out/ is capped at 64 MiB)
for f in $O/${R}_*.ncu-rep; do case $f in *prefill896_attn72*) ;; *) rm -f $f;; esac; done
