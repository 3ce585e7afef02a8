"""ORACLE -- test infrastructure only (see oracle/scmc_oracle.c header)."""
