"""`phifem` import name of the reference, served by the B200-native implementation (phifem_b200)."""
