from oracle.capsule_ref import CapsuleLinear  # noqa: F401
