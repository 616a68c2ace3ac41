#!/bin/bash
# profiles/capture_r02.sh -- round-2 evidence, run under gpurun on ONE GPU:
#   launch lists (ncu --metrics gpu__time_duration.sum, cold-cache and serialised: compare SHARES) of short bench
#   runs of c2 / c3 / c4m / c5, and one `ncu --set full` capture of the kernels named in DESIGN.md.
# Raw reports land in gpurun_out/ (scratch); profiles/ncu_raw_summary.py / stall_regions.py turn them into the
# text summaries committed under profiles/r02_*.txt.
set -u
mkdir -p gpurun_out
for CFG in ${LAUNCH_LISTS-c2 c3 c4m c5}; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
      --log-file gpurun_out/launches_r02_${CFG}.csv python bench.py --config ${CFG} --steps 1 --warmup 1 --profile-only --no-cpu-baseline \
      > gpurun_out/bench_under_ncu_r02_${CFG}.log 2>&1
  python profiles/summarize_launches.py gpurun_out/launches_r02_${CFG}.csv > gpurun_out/r02_launches_${CFG}.txt 2>/dev/null
done
cap() { # tag config kernel-regex skip
  case " ${CAPTURES-all} " in *" all "*|*" $1 "*) ;; *) return;; esac
  ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 \
      -o gpurun_out/prof_r02_$1 -f python bench.py --config $2 --steps 1 --warmup 0 --profile-only --no-cpu-baseline \
      > gpurun_out/ncu_r02_$1.log 2>&1
  python profiles/ncu_raw_summary.py gpurun_out/prof_r02_$1.ncu-rep > gpurun_out/r02_ncu_$1.txt 2>/dev/null
  python profiles/stall_regions.py gpurun_out/prof_r02_$1.ncu-rep >> gpurun_out/r02_ncu_$1.txt 2>/dev/null
}
cap c2_eval_fused c2 "k_sparse_assemble" 3
cap c2_trial c2 "k_trial" 2
cap c2_reduce c2 "k_sparse_grad_reduce" 2
cap c4m_bf_step c4m "k_bf_step" 20       # look-ahead panel step: panel tiles + inverse tile + update of the next panel
cap c4m_bf_gemm c4m "k_bf_gemm" 3        # Schur complement of a level (K = all pivots)
cap c5_potrf_step c5 "k_bf_step" 30
cap c5_syrk c5 "k_dense_syrk_dmma" 1
cap c4m_gather c4m "k_extend_gather" 3
cap c4m_leaf c4m "k_leaf_fronts_mma" 2
cap c3_trial c3 "k_batched_trial" 2
ls -la gpurun_out | grep r02_ | head -40
