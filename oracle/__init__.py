"""CPU oracle (test infrastructure only; see sse_oracle.py header)."""
