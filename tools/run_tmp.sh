python tools/diag_slice.py
