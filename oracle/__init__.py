"""CPU oracle — test infrastructure only (see gnn_oracle.py header)."""
