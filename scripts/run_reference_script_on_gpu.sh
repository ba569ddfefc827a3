#!/bin/bash
# Runs the reference's retrieve/retrieve_faiss.py UNCHANGED on a B200 box with `import faiss` = textreact_b200's shim.
# The reference tree is not part of this repo and does not travel with the gpurun snapshot, so the script file is
# handed to the box inside the command line (base64) and written to /tmp there -- nothing of it enters the repo.
# Run from the build container (where /root/reference exists):   bash scripts/run_reference_script_on_gpu.sh
set -e
SRC=${1:-/root/reference/retrieve/retrieve_faiss.py}
B64=$(base64 -w0 "$SRC")
SUM=$(sha256sum "$SRC" | cut -d' ' -f1)
/usr/local/graft/bin/gpurun --timeout 600 -- "mkdir -p /tmp/ref/retrieve gpurun_out && echo $B64 | base64 -d > /tmp/ref/retrieve/retrieve_faiss.py && echo 'reference script sha256 (must equal $SUM):' \$(sha256sum /tmp/ref/retrieve/retrieve_faiss.py) > gpurun_out/reference_script_on_gpu.log && TRX_REFERENCE_SCRIPT=/tmp/ref/retrieve/retrieve_faiss.py python -m pytest tests/test_gpu_dropin.py -v -k 'unchanged_script or script_flow' 2>&1 | tail -15 >> gpurun_out/reference_script_on_gpu.log; cat gpurun_out/reference_script_on_gpu.log"
