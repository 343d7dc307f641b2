"""TEST INFRASTRUCTURE: CPU oracle for the radian hot path (see oracle/radian_oracle.c)."""
