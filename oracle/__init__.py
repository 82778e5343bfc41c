"""CPU oracle (test infrastructure).  See the module docstrings; never imported by the product."""
