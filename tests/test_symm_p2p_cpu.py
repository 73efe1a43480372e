"""Protocol test of the peer-store transport (dist/transport.py::SymmP2P) without GPUs: several
"ranks" live in one process, symmetric memory is a dictionary of CPU tensors behind fake device
addresses, and ``p2p_put`` / ``p2p_wait`` are emulated with the kernels' semantics (a wait that
would have to spin sets the status word instead)."""
import torch

from ddlbench_b200.dist.transport import SymmP2P, TensorSpec


class FakeFabric:
    def __init__(self):
        self.mem = {}                      # base address -> uint8 tensor
        self.allocs = {}                   # (pair, index) -> [SymmetricBuffer-like per rank]
        self.count = {}
        self.next = 1 << 40

    def alloc_for(self, rank):
        def alloc(pair, numel, dtype):
            k = self.count.get((rank, pair), 0)
            self.count[(rank, pair)] = k + 1
            key = (pair, k)
            if key not in self.allocs:
                bufs = []
                for _ in pair:
                    t = torch.zeros(numel, dtype=dtype)
                    self.mem[self.next] = t.view(torch.uint8)
                    bufs.append((self.next, t))
                    self.next += 1 << 32
                self.allocs[key] = bufs
            bufs = self.allocs[key]

            class Buf:
                tensor = bufs[pair.index(rank)][1]
                ptrs = [b[0] for b in bufs]
            return Buf
        return alloc

    def resolve(self, addr):
        base = max(b for b in self.mem if b <= addr)
        return self.mem[base], addr - base


class FakeP2PExt:
    def __init__(self, fabric):
        self.f = fabric

    @staticmethod
    def raw_bytes(t):
        if t.dim() == 4 and not t.is_contiguous():
            t = t.permute(0, 2, 3, 1)            # channels-last memory order
        return t.contiguous().view(-1).view(torch.uint8)

    def p2p_put(self, src, dst_addr, flag_addr, local_seq, ticket, blocks):
        if src.numel():
            mem, off = self.f.resolve(dst_addr)
            raw = self.raw_bytes(src)
            mem[off:off + raw.numel()] = raw
        local_seq += 1
        mem, off = self.f.resolve(flag_addr)
        mem[off:off + 4].view(torch.int32)[0] = int(local_seq)

    def p2p_wait(self, flag_addr, local_seq, status):
        mem, off = self.f.resolve(flag_addr)
        want = int(local_seq) + 1
        if int(mem[off:off + 4].view(torch.int32)[0]) < want:
            status[0] = 3                         # the kernel would spin here
        local_seq.fill_(want)


def make(nranks, pairs, slot_bytes=4096, ring=2):
    fab = FakeFabric()
    ext = FakeP2PExt(fab)
    return [SymmP2P("cpu", r, pairs, slot_bytes, ring, fab.alloc_for(r), ext) for r in range(nranks)]


SPECS = [TensorSpec("act", (2, 8, 4, 4), torch.bfloat16, True), TensorSpec("target", (2,), torch.int64)]
GSPEC = [TensorSpec("gact", (2, 8, 4, 4), torch.bfloat16, True)]


def msg(i):
    x = torch.full((2, 8, 4, 4), float(i)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x[0, 1, 2, 3] = -i
    return [x, torch.tensor([i, i + 1])]


def test_ring_credits_and_zero_copy_views():
    r0, r1 = make(2, [(0, 1)])
    pend = [r1.post_recv(0, SPECS, "fwd")]
    r0.send(1, msg(1), "fwd")
    a1, t1 = pend[0].wait()
    assert torch.equal(a1, msg(1)[0]) and torch.equal(t1, msg(1)[1])
    assert a1.is_contiguous(memory_format=torch.channels_last)
    pend.append(r1.post_recv(0, SPECS, "fwd"))           # slot 1
    r0.send(1, msg(2), "fwd")
    a2, _ = pend[1].wait()
    assert torch.equal(a2, msg(2)[0]) and torch.equal(a1, msg(1)[0])      # slot 0 still intact
    pend.append(r1.post_recv(0, SPECS, "fwd"))           # slot 0 again: the view a1 is recycled
    r0.send(1, msg(3), "fwd")
    a3, t3 = pend[2].wait()
    assert torch.equal(a3, msg(3)[0]) and torch.equal(t3, msg(3)[1]) and a3.data_ptr() == a1.data_ptr()
    # the reverse channel runs independently on the same pair
    back = r0.post_recv(1, GSPEC, "bwd")
    r1.send(0, [msg(7)[0]], "bwd")
    assert torch.equal(back.wait()[0], msg(7)[0])
    assert int(r0.status) == 0 and int(r1.status) == 0
    assert r0.bytes_sent == r1.bytes_received > 0


def test_send_without_credit_would_block():
    r0, r1 = make(2, [(0, 1)])
    r0.send(1, msg(1), "fwd")                             # the receiver never posted: no credit
    assert int(r0.status) == 3


def test_wait_before_data_would_block():
    r0, r1 = make(2, [(0, 1)])
    p = r1.post_recv(0, SPECS, "fwd")
    p.wait()
    assert int(r1.status) == 3


def test_middle_stage_has_two_independent_peers():
    r0, r1, r2 = make(3, [(0, 1), (1, 2)])
    p01 = r1.post_recv(0, SPECS, "fwd")
    p12 = r2.post_recv(1, SPECS, "fwd")
    r0.send(1, msg(4), "fwd")
    got = p01.wait()
    r1.send(2, [got[0] * 2, got[1]], "fwd")
    a, t = p12.wait()
    assert torch.equal(a, msg(4)[0] * 2) and torch.equal(t, msg(4)[1])
    g21 = r1.post_recv(2, GSPEC, "bwd")
    g10 = r0.post_recv(1, GSPEC, "bwd")
    r2.send(1, [msg(5)[0]], "bwd")
    r1.send(0, [g21.wait()[0] + 1], "bwd")
    assert torch.equal(g10.wait()[0], msg(5)[0] + 1)
    assert all(int(r.status) == 0 for r in (r0, r1, r2))
