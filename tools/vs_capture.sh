cd /root/repo
ncu --set full --clock-control none -k regex:"verify_arith_kernel|verify_pre_kernel" -c 2 -o /tmp/prof_vs -f python tools/op_probe.py verify_shared 65 1048576 1 > /dev/null 2>&1
python tools/summarize_ncu.py round2/verify_shared65 --rep /tmp/prof_vs.ncu-rep --lanes 1048576 --ps 65 --note "\`tools/op_probe.py verify_shared 65 1048576 1\`: verify_pre and verify_arith of a verify under ONE shared public key (cached A-hat, tr, t1-hat: BASELINE configs[2]'s shape) at 2^20 lanes, ML-DSA-65."
mkdir -p gpurun_out/profiles_round2
cp profiles/round2/verify_shared65.md profiles/ncu_kernels.json gpurun_out/profiles_round2/
