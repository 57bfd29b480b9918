# Captures of the round-2/session-3 kernels on the GPU box: bash profiles/capture_round3.sh  (outputs under gpurun_out/)
set -x
python profiles/oneshot_probe.py 8 > gpurun_out/r3l_oneshot.log 2>&1
for k in lz_find lz_prev lz_stitch xhuff_encode_exact; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/r3l_xh_$k python profiles/run_kernel.py xpress_huff 4096 1 > gpurun_out/r3l_ncu_$k.log 2>&1
done
for k in lz_find xpress_emit_exact; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/r3l_xp_$k python profiles/run_kernel.py xpress 256 1 > gpurun_out/r3l_ncu_xp_$k.log 2>&1
done
compute-sanitizer --tool memcheck python tests/sanitize_small.py > gpurun_out/r3l_memcheck.log 2>&1
tail -3 gpurun_out/r3l_memcheck.log
cat gpurun_out/r3l_oneshot.log
