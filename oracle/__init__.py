"""ORACLE -- test infrastructure only (never imported by mc3_b200/).

A CPU restatement of the reference's sampling hot path (pcubillos/mc3 3.2.1):

  oracle_kernels.c   plain-C chisq / residuals / priors / dwt_chisq / daub4 /
                     binrms (+invgamma) / binarray          -> liboracle.so
  kernels.py         ctypes + numpy wrappers with the mc3.stats signatures
  mcmc.py            numpy restatement of mc3/chain.py + mc3/mcmc_driver.py
                     (single-process semantics) with a draw recorder
  ref.py             loaders for oracle/_ref (the reference's own C extensions,
                     compiled by oracle/Makefile) and -- in the authoring
                     container only -- for the reference's Python modules
  make_golden.py     writes tests/golden/*.npz from the real reference

Parity status: pinned (see tests/test_oracle.py).
Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg and
--impl reference) may import this package.
"""
