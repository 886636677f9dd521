"""Empty pyplot shim."""
