"""CPU checkers for the B200 codec.  TEST INFRASTRUCTURE ONLY (see oracle/huf_oracle.c)."""
