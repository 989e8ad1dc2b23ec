"""Test infrastructure: CPU oracle of the FacialMMT forward path. Not product code (see facialmmt_oracle.py)."""
