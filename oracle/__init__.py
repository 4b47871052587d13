"""CPU oracle of the qLDPCsim decoder path -- test infrastructure only (see oracle/oracle.py)."""
