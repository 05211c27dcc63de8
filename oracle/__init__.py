"""CPU oracle for the correlation hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package (understanding_flow_robustness_b200) never does;
tests/test_no_oracle_in_product.py enforces that.
"""
