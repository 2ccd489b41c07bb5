# end-of-round sequence on one B200 (gpurun -- 'bash tools/gpu/final_all.sh'): the whole GPU suite, the ncu captures of the
# current sources, then -- with the fresh ncu_metrics.json in place -- the final bench lines of every workload
timeout 1700 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/tests.log; tail -4 gpurun_out/tests.log
bash tools/gpu/profile_r02.sh > gpurun_out/profile_r02.log 2>&1; tail -2 gpurun_out/ncu_extract.log
cp gpurun_out/ncu_metrics.json profiles/ncu_metrics.json
bash tools/gpu/final_sequence.sh
