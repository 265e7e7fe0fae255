"""CPU oracle (test infrastructure only) -- see oracle/tscnet_oracle.py."""
