"""TEST INFRASTRUCTURE: CPU oracle for the stereo-matching hot path (see stereo_oracle.c)."""
