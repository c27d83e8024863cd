#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name: $*"; "$@" > gpurun_out/segv_$name.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/segv_$name.log | cut -c1-200; }
run D python -X faulthandler -m pytest tests/test_reference_binding_gpu.py -m gpu -q
run A env PYTHONMALLOC=debug python -X faulthandler -m pytest tests/test_reference_binding_gpu.py -m gpu -q
run B python -X faulthandler -m pytest tests -m gpu -q -k "ref_parity"
run C python -X faulthandler -m pytest tests -m gpu -q -k "gradient"
run E env PYTHONMALLOC=debug python -X faulthandler -m pytest tests -m gpu -q -k "optim or gradient or embedding or ref_parity or reference_binding"
run F python -X faulthandler -m pytest tests -m gpu -q -k "gradient or ref_parity or reference_binding"
