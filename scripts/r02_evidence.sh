#!/bin/bash
# Round-2 evidence set (run on the GPU box through gpurun): bench lines of every workload, the ncu launch list of the default
# bench command, DRAM bytes / issue activity of the dominant kernel of each workload.  Output: gpurun_out/r02_*.
O=gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio
python bench.py > $O/r02_bench_c3.json 2> $O/r02_bench_c3.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference.json 2>/dev/null
for w in c4 c5 c2 market gym c1; do python bench.py --workload $w > $O/r02_bench_$w.json 2>/dev/null; done
python bench.py --workload c5 --envs 512 --no-cpu > $O/r02_bench_c5_512.json 2>/dev/null
python bench.py --workload c5 --envs 256 --no-cpu > $O/r02_bench_c5_256.json 2>/dev/null
python bench.py --workload c2 --envs 512 --no-cpu > $O/r02_bench_c2_512books.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-logs > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:k_sim -c 1 --csv --log-file $O/r02_c3_ksim_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-logs --no-secondary > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:k_sim -c 1 --csv --log-file $O/r02_c4_ksim_dram.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:k_deepw -s 1 -c 1 --csv --log-file $O/r02_c5_kdeepw_dram.csv python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:k_apply -c 1 --csv --log-file $O/r02_c2_kapply_dram.csv python bench.py --workload c2 --envs 512 --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:k_sim -c 1 --csv --log-file $O/r02_market_ksim_dram.csv python bench.py --workload market --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:k_apply -s 8 -c 1 --csv --log-file $O/r02_gym_kapply_dram.csv python bench.py --workload gym --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la $O | grep r02_
