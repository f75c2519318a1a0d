"""Drop-in `modules` package: same module/class names as the reference's
modules/ directory, computing on the sm_100a kernels behind include/tgn_b200.h."""
