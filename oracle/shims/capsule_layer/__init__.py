"""Stand-in for the un-vendored third-party `capsule_layer` package (see oracle/capsule_ref.py)."""
