#!/bin/bash
bash tools/gpu_runs/gpu_run14.sh
bash tools/gpu_runs/gpu_run13_2gpu.sh 2
