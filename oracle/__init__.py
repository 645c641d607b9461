"""CPU oracle - TEST INFRASTRUCTURE ONLY (see qutip4_oracle.py).  The product never imports this."""
