#!/bin/bash
# ncu captures of the bench command (one gpurun call: all ncu runs of a call count as one).  Outputs under gpurun_out/.
# ONE pass in flight (--lanes 1): with several lane threads launching graphs concurrently the process dies inside ncu (SIGSEGV; round 1
# saw glibc heap-corruption aborts in the same situation) while the same command without ncu, and under TSan / ASan, is clean.
cd ${GRAFT_REPO_ROOT:-.}
CMD="python -X faulthandler bench.py --steps 20 --warmup 3 --lanes ${PROF_LANES:-1} --queue-lanes 1 --submitters 1 --extras 0"
K='regex:k_replay|k_decompress|k_vprep|k_msm|k_encode|k_scan'
$CMD > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k "$K" -s 200 -c 660 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu1.log 2>&1
echo "launch list rc=$?"; tail -25 gpurun_out/r2_ncu1.log | cut -c1-300
$CMD > gpurun_out/r2_prof_plain2.json 2> gpurun_out/r2_prof_plain2.err && \
ncu --set full --clock-control none --import-source on -k 'regex:k_msm_bucket_thread|k_decompress_proofs|k_replay_sm|k_vprep_vector|k_msm_reduce|k_vprep_proof|k_msm_digits|k_vprep_weight' -s 160 -c 12 -o gpurun_out/r2_prof_full $CMD > gpurun_out/r2_ncu2.log 2>&1
echo "set full rc=$?"; tail -5 gpurun_out/r2_ncu2.log | cut -c1-300; ls -la gpurun_out/r2_prof_full.ncu-rep gpurun_out/r2_launches.csv
# gpurun brings back at most 64 MiB: export what the summaries need as CSV here, keep the report only while it fits
ncu -i gpurun_out/r2_prof_full.ncu-rep --page raw --csv > gpurun_out/r2_prof_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_full.ncu-rep --page source --csv -k regex:k_msm_bucket_thread -c 1 > gpurun_out/r2_prof_bucket_source.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_full.ncu-rep --page details --csv > gpurun_out/r2_prof_full_details.csv 2>/dev/null
if [ "${PROF_PROVER:-0}" = "1" ]; then
# prover: the fixed-base sum kernel of one 1024-proof call
PCMD="python scripts/prove_lanes_probe.py 1024 1"
$PCMD > gpurun_out/r2_prof_prove_plain.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_fb_msm|k_encode|k_prove_round_pre_fb' -s 40 -c 6 -o gpurun_out/r2_prof_prove $PCMD > gpurun_out/r2_ncu3.log 2>&1
echo "prove set full rc=$?"; tail -3 gpurun_out/r2_ncu3.log | cut -c1-300
ncu -i gpurun_out/r2_prof_prove.ncu-rep --page raw --csv > gpurun_out/r2_prof_prove_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_prove.ncu-rep --page source --csv -k regex:k_fb_msm -c 1 > gpurun_out/r2_prof_fb_source.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_prove.ncu-rep --page details --csv > gpurun_out/r2_prof_prove_details.csv 2>/dev/null
ls -la gpurun_out/; du -sm gpurun_out
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/r2_prof_prove.ncu-rep; fi
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/r2_prof_full.ncu-rep; fi
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 150 --csv --log-file gpurun_out/r2_launches_prove.csv $PCMD > gpurun_out/r2_ncu4.log 2>&1
echo "prove launch list rc=$?"
du -sm gpurun_out
fi
