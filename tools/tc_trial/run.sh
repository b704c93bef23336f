# tensor-core trial on a B200: tools/tc_trial/run.sh <tag>  -> gpurun_out/<tag>_tc_trial.json, <tag>_tc_*.ncu-rep
T=${1:-r2}
mkdir -p gpurun_out
timeout 120 tools/tc_trial/tc_matrix 2000 > gpurun_out/${T}_tc_trial.json 2> gpurun_out/${T}_tc_trial.err; echo "rc=$?"; cat gpurun_out/${T}_tc_trial.json; tail -3 gpurun_out/${T}_tc_trial.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stage_ -s 2 -c 2 -o gpurun_out/${T}_tc_trial -f tools/tc_trial/tc_matrix 300 > /dev/null 2>&1
ls -la gpurun_out | grep ${T}_tc
