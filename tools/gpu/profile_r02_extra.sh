# extra ncu captures of round 2: the fused TopoNet kernel and the graph-stage kernels of one C2 scene
# (gpurun -- 'bash tools/gpu/profile_r02_extra.sh'); summaries -> gpurun_out/r02_topo_top.txt, r02_graph_top.txt
set -x
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scene"
O=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:toponet_tc_kernel -s 1 -c 1 -o $O/r02_topo $B > $O/ncu_topo.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"nms_round_kernel|knn_kernel|aggregate_warp_kernel|fuse_masks_kernel" -c 16 -o $O/r02_graph python tools/scene_once.py numpy > $O/ncu_graph.log 2>&1
for r in r02_topo r02_graph; do python tools/ncu_top.py $O/$r.ncu-rep 12 > $O/${r}_top.txt 2>&1; done
ls -la $O/r02_topo.ncu-rep $O/r02_graph.ncu-rep
