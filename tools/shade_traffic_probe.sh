#!/bin/bash
# DRAM bytes of the first k_shade launches of the material grid in the unsorted and the material-sorted order (8 spp: 16.6 M vertices in round 0)
mkdir -p gpurun_out
for m in 0 1; do
timeout -k 10 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum --clock-control none -k regex:'k_(shade|bin)' -s 6 -c 6 --csv --log-file gpurun_out/shade_probe_$m.csv python bench.py --workload material_grid --steps 1 --warmup 1 --spp 8 --no-e2e --no-cpu-baseline --no-sub --opt sort_by_material=$m > gpurun_out/shade_probe_$m.log 2>&1; echo "mode $m rc=$?"
done
