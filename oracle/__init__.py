"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU (torch fp32) restatement of the reference's G+D training-step arithmetic
(/root/reference model.py + cub_trainer_splitz_cap_ca.py + trainer.py).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product path (text2img_ekl_b200/) never does.

Parity status: pinned against outputs of the real reference run in the build
container (oracle/gen_golden.py -> tests/golden/*.npz) for everything except
the third-party `capsule_layer.CapsuleLinear` arithmetic, which is absent from
/root/reference and un-pinned upstream ("parity unpinned" -- see
oracle/capsule_ref.py and DESIGN.md).
"""
