#!/bin/sh
# Installs the UNMODIFIED reference package into baseline/_ref (git-ignored, travels to the GPU box).
# --no-deps: the reference pins numpy==1.26.3 / tqdm / pandas versions the offline wheelhouse does not
# hold; the image's numpy 2.3, tqdm and pandas are used instead (no behavioural difference
# found on this path).  mpi4py is not installable here: oracle/mpi4py_stub stands in for it.
set -e
cd "$(dirname "$0")/.."
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target baseline/_ref --upgrade /root/reference
