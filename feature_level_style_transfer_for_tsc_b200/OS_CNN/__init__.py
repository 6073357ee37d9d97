"""Drop-in mirror of the reference package ``OS_CNN`` (module paths ``OS_CNN.OS_CNN`` and
``OS_CNN.OS_CNN_Structure_build``), backed by the sm_100a kernels of libtsc_b200."""
