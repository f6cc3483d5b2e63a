#!/bin/bash
# Produce the round's measurement artefacts on the GPU box into gpurun_out/<tag>_* (summarised into profiles/ by
# scripts/summarize_profiles.py):
#   the ncu launch list of the bench command (short burn-in), one `ncu --set full` capture of the likelihood kernel
#   (raw metrics exported as CSV, the .ncu-rep itself kept), compute-sanitizer memcheck + racecheck of smoke().
# usage: scripts/capture_profiles.sh <tag>
tag=${1:-r02}
o=gpurun_out
mkdir -p $o
B="python bench.py --steps 4 --warmup 3 --burn 20 --no-cpu-baseline --legs none"
# launch list: same command as the bench, short (sweeps are replayed from CUDA graphs: ncu profiles the kernel nodes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $o/${tag}_launches.csv \
    $B > $o/${tag}_ncu_bench.log 2>&1
# full capture of the dominant kernel: a launch of the timed region (1 init + 2 x 20 burn-in launches precede it)
X=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_fp64.sum
ncu --set full --metrics $X --import-source on --clock-control none -k regex:logl_rv -s 46 -c 1 -o $o/${tag}_logl -f \
    $B > $o/${tag}_ncu_full.log 2>&1
ncu -i $o/${tag}_logl.ncu-rep --page raw --csv > $o/${tag}_logl_raw.csv 2>/dev/null
compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_sanitizer.log 2>&1
compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" >> $o/${tag}_sanitizer.log 2>&1
grep "ERROR SUMMARY" $o/${tag}_sanitizer.log
tail -2 $o/${tag}_sanitizer.log
ls -la $o/${tag}_*
