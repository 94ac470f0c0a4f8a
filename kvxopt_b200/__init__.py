"""kvxopt_b200: B200-native sparse direct solvers behind kvxopt's cholmod / klu interface."""
