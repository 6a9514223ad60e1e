"""`env` package shim: lets the reference's scripts import the B200 implementation unchanged."""
