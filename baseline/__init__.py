"""Access to the UNMODIFIED reference (xiahouzuoxin/torchctr) installed under ``baseline/_ref`` -- test / bench
infrastructure only; nothing in ``torchctr_b200`` imports it."""
