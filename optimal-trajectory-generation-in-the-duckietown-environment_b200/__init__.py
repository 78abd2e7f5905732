"""B200-native Frenet lattice planner hot path."""
