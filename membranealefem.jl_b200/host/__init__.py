"""Host-side mirror of the reference's Input / Analysis API for the path that calls calc_r_K."""
