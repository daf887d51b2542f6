"""Test suite of gpflowpilco_b200: CPU tests (oracle, host logic, C ABI) and `-m gpu` parity tests (see tests/conftest.py)."""
