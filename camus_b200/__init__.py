"""camus_b200: B200-native hot path of CAMUS (quartet counting / filtering / edge scoring).

  camus_b200.cabi     -- ctypes binding of the CUDA C-ABI (include/camus_b200.h)
  camus_b200.hostlib  -- ctypes binding of the C++ host mirror (parse, flatten, DP, network)
  camus_b200.build    -- in-tree builds of both libraries and the `camus` CLI
"""
__version__ = "0.1.0"
