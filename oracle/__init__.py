"""CPU oracle of the reference's trust-region DP -- TEST INFRASTRUCTURE ONLY (see bellman_oracle.c)."""
