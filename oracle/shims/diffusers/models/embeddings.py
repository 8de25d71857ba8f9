from oracle.flux_oracle import apply_rotary_emb  # noqa: F401  (block.py:75)
