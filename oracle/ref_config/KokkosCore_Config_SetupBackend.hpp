// hand-written (see KokkosCore_config.h): no device backend => nothing to set up
