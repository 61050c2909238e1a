"""phifem_b200: B200-native level-set cut-cell classification + phi-FEM CSR assembly."""
