"""Drop-in import names of the reference (``nf.flows``, ``nf.models``, ``nf.utils``, ``nf.hmc``)
backed by normalizingflow_b200 — see that package for the implementation."""
