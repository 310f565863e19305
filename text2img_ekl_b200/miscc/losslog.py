"""Scalar logging of the training loop without host synchronisation.

The reference writes tensorboard scalars every 100 iterations from inside train_joint_Dnet (`summary.scalar('D_loss%d'
% idx, errD.data[0])`, cub_trainer_splitz_cap_ca.py:457-460): `.data[0]` / `.item()` blocks the host until the device
has drained, once per discriminator.  Here the loss tensors of a logged iteration are copied to a pinned host slot with
a non-blocking D2H copy followed by an event; the values are read (and written out as JSON lines, or handed to a
tensorboardX-style writer) only once the event has completed -- polled at later iterations, never waited for inside
the loop.
"""
import json
import os

import torch


class AsyncLossLog:
    def __init__(self, path=None, every=100, slots=8, writer=None, width=32):
        self.path, self.every, self.writer = path, every, writer
        self.slots = [dict(buf=None, ev=None, tag=None, busy=False) for _ in range(slots)]
        self.width = width
        self.records = []              # (count, {name: value}) in completion order (also kept for tests / epoch prints)
        self.dropped = 0
        if path:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)

    def due(self, count):
        return self.every > 0 and count % self.every == 0

    def push(self, count, names, values):
        """names: list of scalar names; values: 1-D device (or host) tensor of the same length.  Never blocks: when every
        slot is still in flight the sample is dropped (and counted)."""
        self.poll()
        slot = next((s for s in self.slots if not s["busy"]), None)
        if slot is None:
            self.dropped += 1
            return False
        v = values.detach().reshape(-1).float()
        n = v.numel()
        assert n == len(names) and n <= self.width
        if v.is_cuda:
            if slot["buf"] is None:
                slot["buf"] = torch.zeros(self.width, pin_memory=True)
                slot["ev"] = torch.cuda.Event()
            slot["buf"][:n].copy_(v, non_blocking=True)
            slot["ev"].record()
        else:
            slot["buf"] = torch.zeros(self.width) if slot["buf"] is None else slot["buf"]
            slot["buf"][:n].copy_(v)
        slot["tag"], slot["busy"] = (int(count), list(names), v.is_cuda), True
        return True

    def poll(self, wait=False):
        """Consume every slot whose copy has landed (wait=True: all of them, synchronising on their events)."""
        done = 0
        for s in self.slots:
            if not s["busy"]:
                continue
            count, names, on_dev = s["tag"]
            if on_dev:
                if wait:
                    s["ev"].synchronize()
                elif not s["ev"].query():
                    continue
            rec = {k: float(x) for k, x in zip(names, s["buf"][:len(names)].tolist())}
            s["busy"] = False
            self.records.append((count, rec))
            done += 1
            if self.writer is not None:                    # tensorboardX-style: add_scalar(name, value, step)
                for k, x in rec.items():
                    self.writer.add_scalar(k, x, count)
            if self.path:
                with open(self.path, "a") as f:
                    f.write(json.dumps({"count": count, **rec}) + "\n")
        return done

    def flush(self):
        return self.poll(wait=True)
