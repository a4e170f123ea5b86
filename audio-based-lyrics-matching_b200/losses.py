"""placeholder -- filled in with the fused NT-Xent / CLEWS modules (build order: eval path first)."""
