#!/bin/bash
bash tools/final_validate.sh r2s
bash tools/gpu_sanitize.sh r2s 2>&1 | tail -14
