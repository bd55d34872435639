"""rapidcfd-dev_b200 -- Blackwell-native lduMatrix linear-solver core.

The directory name carries a hyphen (it mirrors the reference repo's name); import it
with importlib.import_module("rapidcfd-dev_b200") or through tests/conftest.py's `pkg`.
"""
