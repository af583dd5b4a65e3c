"""CPU checkers for the QGT hot path.  TEST INFRASTRUCTURE ONLY — never imported by the product."""
