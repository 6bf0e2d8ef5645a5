"""Stand-in for mpi4py (TEST INFRASTRUCTURE ONLY — never imported by numpitron_b200).

The image has no MPI library, no mpirun and no mpi4py, so the unmodified
reference (`/root/reference/src/numpitron`) cannot be imported as-is.  This
package provides exactly the `mpi4py.MPI` surface the reference touches
 so it can be run as the parity oracle:

* world size 1: every collective is the identity / a local copy;
* world size > 1 (env WORLD_SIZE/RANK set, one OS process per rank): the
  collectives are carried by `torch.distributed` with the gloo backend.
"""
from . import MPI  # noqa: F401
