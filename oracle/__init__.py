"""TEST INFRASTRUCTURE ONLY: CPU oracle of gpupsat_b200 (see oracle/gpsat_oracle.cpp, oracle/ref/)."""
