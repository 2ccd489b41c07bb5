# final-build sanity on 2 GPUs: smoke(), the reference arm, the NCCL scene test, the N=2 bench line
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/ref_arm.err | tail -1 | cut -c1-400
bash tools/gpu/multi_gpu_check.sh 2
