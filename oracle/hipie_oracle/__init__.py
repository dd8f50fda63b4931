"""CPU oracle of HIPIE's inference hot path — test infrastructure, not product code (see oracle/README.md)."""
