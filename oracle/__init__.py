"""CPU oracle (test infrastructure only — see oracle/gas_oracle.h)."""
