"""CPU oracle — test infrastructure only (see crnn_oracle.c)."""
