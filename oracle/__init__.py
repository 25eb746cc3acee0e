"""CPU oracle (TEST INFRASTRUCTURE ONLY — see oracle/oracle.cpp header). Never imported by the product."""
