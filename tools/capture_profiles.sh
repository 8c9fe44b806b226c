#!/bin/bash
# Round-2 ncu evidence, one GPU (run under gpurun): launch list of a short bench command + one full capture per kernel.
# Raw pages land in gpurun_out/ as CSV; copy the ones to keep into profiles/.
set -u
O=gpurun_out
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --streams 4096 --no-extras --no-roofline --no-cpu > $O/r02_launches_bench.log 2>&1
cap() {  # name, kernel regex, launches to skip, command...
    local name=$1 regex=$2 skip=$3
    shift 3
    timeout 600 $NCU --set full --import-source on -k regex:$regex -s $skip -c 1 -f -o $O/$name "$@" > $O/$name.log 2>&1
    ncu -i $O/$name.ncu-rep --page raw --csv > $O/r02_ncu_${name}_raw.csv 2>/dev/null
    python tools/ncu_digest.py $O/$name.ncu-rep 25 > $O/r02_ncu_${name}_digest.txt 2>&1
    rm -f $O/$name.ncu-rep
}
cap k2p dtw_windows_d16 3 python tools/prof_pipeline.py 2048 5 1
cap k1 mfcc_frames2 2 python tools/prof_pipeline.py 2048 5 1
cap k2c dtw_windows_cadence 40 python tools/prof_cadence.py 4096 5 125
cap k2s dtw_pairs_stream4 1 python tools/prof_dtw_stream.py 1000000
ls -la $O | tail -20
