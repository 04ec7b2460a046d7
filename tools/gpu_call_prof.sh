#!/bin/bash
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1c_train_tf32_launches.csv python tools/train_steps_tf32.py 3 > gpurun_out/ncu_train.log 2>&1
tail -3 gpurun_out/ncu_train.log; wc -l gpurun_out/r1c_train_tf32_launches.csv
