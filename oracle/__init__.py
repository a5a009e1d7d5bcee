"""CPU oracle for the whole-body QP hot path.  TEST INFRASTRUCTURE (see qppvm_oracle.c)."""
