"""CPU oracle of the SUMMA GEMM hot path — TEST INFRASTRUCTURE ONLY (see oracle/gemm_oracle.c)."""
