"""psc_b200: B200-native (sm_100a) implementation of psc-code/psc's per-timestep
particle-in-cell hot path, behind PSC's PscConfig plugin surface.

The product is the CUDA library psc_b200/lib/libpsc_b200.so (C ABI in
include/psc_b200.h) plus the C++ wrapper types in include/psc_b200/.  This Python
package is the host-side mirror of the same interface for tests and benchmarks."""
from ._lib import GridDesc, StepParams, CollisionParams, HeatingParams, JPath, PscB200Error, load, check  # noqa: F401
from .api import (Moment, MOMENT_N, MOMENT_V, MOMENT_P, MOMENT_T, MOMENT_ALL, MOMENT_RHO_NC,  # noqa: F401
                  Grid, Mparticles, Mfields, MfieldsState, energies, write_checkpoint, read_checkpoint, PushParticles, Sort, BndParticles,  # noqa: F401
                  Bnd, BndFields, PushFields, Marder, Checks, Psc, Collision, Heating, BoundaryInjector, ParticleGeneratorMaxwellian, PRT_DTYPE,
                  ItemJeh, OutputFieldItemParams, WriterMemory, OutputFieldsItem, OutputFields, OutputMoments,
                  JXI, JYI, JZI, EX, EY, EZ, HX, HY, HZ, NR_FIELDS,
                  BND_FLD_OPEN, BND_FLD_PERIODIC, BND_FLD_CONDUCTING_WALL, BND_FLD_ABSORBING,
                  BND_PRT_REFLECTING, BND_PRT_PERIODIC, BND_PRT_ABSORBING, BND_PRT_OPEN,
                  DEPOSIT_DEFAULT, DEPOSIT_VAR1, DEPOSIT_SPLIT)
