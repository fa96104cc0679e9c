#!/bin/bash
# True reference baseline: times the UNMODIFIED reference (TensorFlow 2.x + Keras 2.3.1) on the
# host cores -- generation (BrainGenerator.generate_brain) and a 5-step training() -- wherever that stack exists.
#   bench_ref/run_reference_tf.sh /path/to/SynthSR-reference [threads]
# In this project's image TensorFlow is not installable: the script then prints "not run" and exits 0; the number
# reported by bench.py's cpu_baseline is the NumPy / torch-CPU restatement and is labelled as such, never as TF.
REF=${1:-/root/reference}
THREADS=${2:-$(nproc)}
HERE=$(cd "$(dirname "$0")" && pwd)
if ! python -c "import tensorflow, keras, nibabel" 2>/dev/null; then
  echo '{"reference_tf_cpu": "not run (TensorFlow / Keras / nibabel unavailable)"}'
  exit 0
fi
CUDA_VISIBLE_DEVICES=-1 python "$HERE/reference_tf_timing.py" "$REF" "$THREADS"
