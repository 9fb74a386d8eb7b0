#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "sharded_fast_path" 2>&1 | grep -E "^E  .*(assert|Error|error)|passed|failed|FAILED" | head -40
