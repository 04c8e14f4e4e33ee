#!/bin/bash
# The ncu evidence of a round, in one session (B200_PROFILING.md recipe): the launch list of the bench command and
# --set full captures of every hot kernel at 2^20 lanes, summarised ON THE BOX (tools/summarize_ncu.py) so that only the
# summaries and one report travel back: gpurun_out/profiles_round2/{sign65.md, ops65.md, pipeline87.md, ncu_kernels.json}.
cd "$(dirname "$0")/.."
TAG=round2
OUT=gpurun_out/profiles_$TAG
mkdir -p $OUT profiles/$TAG
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-verify-leg --no-cpu-baseline --no-configs > $OUT/launches_bench.json 2> /dev/null
ncu --set full --clock-control none --import-source on -k regex:"expand_a_kernel|sign_key_kernel|sign_prep_kernel|mask_kernel|sign_w_staged|ctilde_kernel|sign_z_kernel" \
    -c 7 -o /tmp/prof_sign -f python tools/op_probe.py sign 65 1048576 1 > /dev/null 2>&1
python tools/summarize_ncu.py $TAG/sign65 --launches $OUT/launches.csv --rep /tmp/prof_sign.ncu-rep --lanes 1048576 --ps 65 \
    --note "Launch list: \`python bench.py --steps 2 --warmup 1 --no-verify-leg --no-cpu-baseline --no-configs\` (3 warm-up + 2 timed device-resident steps of 2^20 lanes, then the end-to-end legs through the chunk pipelines).  Full captures: \`tools/op_probe.py sign 65 1048576 1\`, the first launch of each kernel = the first round, all 2^20 lanes live, two candidates per lane.  ML-DSA-65 sign, distinct keys."
ncu --set full --clock-control none -k regex:"kg_fused|rej_bounded|shake_rows|verify_fused|verify_pre|ntt_kernel|kg_seed" \
    -c 16 -o /tmp/prof_ops -f python tools/bench_ops.py --sets 65 --reps 1 > /dev/null 2>&1
python tools/summarize_ncu.py $TAG/ops65 --rep /tmp/prof_ops.ncu-rep --lanes 1048576 --ps 65 \
    --note "\`tools/bench_ops.py --sets 65 --reps 1\` under \`ncu --set full --clock-control none\`: the NTT / INTT microbench (configs[1]), keygen (ExpandA fused with the matrix-vector product: kg_fused) and per-lane-key verify (verify_fused) at 2^20 lanes, ML-DSA-65; every kernel appears twice (warm-up + timed call)."
ncu --set full --clock-control none -k regex:"verify_arith_kernel|verify_pre_kernel" \
    -c 2 -o /tmp/prof_vs -f python tools/op_probe.py verify_shared 65 1048576 1 > /dev/null 2>&1
python tools/summarize_ncu.py $TAG/verify_shared65 --rep /tmp/prof_vs.ncu-rep \
    --note "\`tools/op_probe.py verify_shared 65 1048576 1\`: verify_pre and verify_arith of a verify under ONE shared public key (cached A-hat, tr, t1-hat: BASELINE configs[2]'s shape) at 2^20 lanes, ML-DSA-65."
ncu --set full --clock-control none -k regex:"kg_arith_kernel|verify_arith_kernel|t1hat|expand_a_kernel" \
    -c 4 -o /tmp/prof_pipe -f python tools/pipeline_probe.py 262144 1 87 > /dev/null 2>&1
python tools/summarize_ncu.py $TAG/pipeline87 --rep /tmp/prof_pipe.ncu-rep --lanes 32768 --ps 87 \
    --note "\`tools/pipeline_probe.py 262144 1 87\` (mldsa_pipeline_batch, chunks of 32 768 lanes): the kernels only the A-hat-keeping paths use -- expand_a + kg_arith (keygen into key sets / the chained pipeline) and verify_arith on cached A-hat."
cp profiles/$TAG/sign65.md profiles/$TAG/ops65.md profiles/$TAG/pipeline87.md profiles/$TAG/verify_shared65.md profiles/ncu_kernels.json $OUT/
cp /tmp/prof_sign.ncu-rep $OUT/prof_sign65.ncu-rep
ls -la $OUT
