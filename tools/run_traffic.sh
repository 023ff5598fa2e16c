for w in cornell material_grid terrain; do
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_(shade|extend|shadow)' --csv --log-file gpurun_out/traffic_$w.csv python bench.py --workload $w --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-sub > gpurun_out/traffic_$w.log 2>&1
done
ls -la gpurun_out/traffic_*
