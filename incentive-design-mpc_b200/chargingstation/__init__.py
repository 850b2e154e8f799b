"""Drop-in mirror of the reference's ``chargingstation`` package for the
lower-level MPC hot path, backed by hand-written sm_100a CUDA kernels behind a
C ABI (``include/lompc_b200.h``).  Put ``incentive-design-mpc_b200/`` on
``PYTHONPATH`` exactly as the reference asks for its own root (README.md:27-31);
``from chargingstation.lompc import LoMPC, LoMPCConstants`` then resolves here.
There is no CPU fallback: without the built library or without a CUDA device
the solve calls raise."""
