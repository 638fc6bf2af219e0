"""Batched early-exit inference / evaluation engine — the public entry point the benchmark times.

It performs, for a whole batch on the device, what the reference does one image at a time in
`br_evaluator` (eval_br_ent.py:51-70) and `eval_ee_deeplabv3.__call__` (ee_dnn_op_ne.py:51-108):

    sections -> exit head -> (fused up-sample + softmax + entropy + argmax) -> per-image decision

with the per-image rule of the reference (first early exit i >= skip whose mean normalised entropy
is < tau, else the final exit), plus the integer confusion matrices of the exit taken. Nothing but
the small per-image results ever leaves the device, and full-resolution logits are never written.

`skip_compute=True` adds what the reference only accounts for (it always runs the tail,
ee_dnn_op_ne.py:91-101): after each gate the still-active images are compacted (device-side list
from `eeseg_exit_gate_decide`) and only they run through the next backbone section.
"""
import torch

from . import ops


class EarlyExitEngine:
    def __init__(self, net, n_classes, tau, metric='ent', size=1, skip=0, skip_compute=False,
                 use_graph=False, input_dtype=torch.float32, target_dtype=torch.int64):
        """use_graph: capture the whole (static) step — backbone sections, heads, gates, decision,
        histogram — into one CUDA graph per input shape and replay it; removes the per-launch host
        overhead of the ~700 launches of a step. With skip_compute the step is captured as one
        graph per (exit stage, number of still-active images) instead: after each gate the host reads
        the 4-byte active count and replays the next stage's graph of that size (_skip_state)."""
        # dtypes of the graph-mode input buffers = what crosses PCIe in evaluate_pipelined: images fp32 (the reference's
        # loader output), bf16 (bit-identical results: the stem rounds to bf16 first; half the bytes) or uint8 (raw
        # pixels, normalised inside the stem kernel with net.input_norm); labels int64 (reference) or uint8 (>= C void)
        assert input_dtype in (torch.float32, torch.bfloat16, torch.uint8) and target_dtype in (torch.int64, torch.uint8)
        self.input_dtype, self.target_dtype = input_dtype, target_dtype
        self.use_graph = use_graph
        self.overlap_gates = True      # early-exit gates on a side stream, overlapping the next section
        self.overlap_heads = True      # ... and the early exits' heads with them (tails of one stream's launches fill with the other's)
        self._side = None
        self._graphs = {}
        self.net = net
        self.C = n_classes
        self.tau = float(tau)
        self.metric = metric.lower()
        assert self.metric in ('ent', 'max', 'min')
        self.size = size if self.metric != 'ent' else 1
        self.skip = skip
        self.skip_compute = skip_compute
        self.E = net.n_branches + 1
        dev = next(net.parameters()).device
        self.device = dev
        # [E+1, C+1, C]: one matrix per exit, last = global (exit actually taken)
        self.cm = torch.zeros((self.E + 1, n_classes + 1, n_classes), dtype=torch.int64, device=dev)
        self.counts = torch.zeros((self.E + 1,), dtype=torch.int64, device=dev)
        self.exited_px = torch.zeros((self.E,), dtype=torch.int64, device=dev)

    def _weights_token(self):
        f = getattr(self.net, 'weights_token', None)
        return f() if f is not None else getattr(self.net, 'weights_epoch', 0)

    def _drop_stale_graphs(self):
        """Graphs captured from an older set of weights are released, not kept: the model's weights_token() (epoch +
        version counters of every parameter and buffer) is compared before every capture lookup AND every replay."""
        token = self._weights_token()
        if getattr(self, '_graph_token', token) != token:
            self._graphs.clear()
        self._graph_token = token
        return token

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def reset(self):
        self.cm.zero_(); self.counts.zero_(); self.exited_px.zero_()

    # ------------------------------------------------------------------------------------------
    def _gate(self, low, out_hw, want_score):
        pool = self.metric != 'ent'
        res = ops.exit_gate(low, out_hw, layout='NHWC', n_classes=self.C, tau=self.tau,
                            want_ent=pool and want_score, want_amax=True,
                            want_score=want_score and not pool)
        if pool and want_score:
            res.score = ops.entropy_pool_mean(res.ent, self.size, self.metric == 'min')
        return res

    def static_inputs(self, shape, with_targets=True):
        """Graph mode: the device buffers the captured step reads. Fill them (e.g. H2D copies
        straight from pinned memory) and call replay() to avoid an extra device copy."""
        g = self._slot(tuple(shape), with_targets)
        return g['X'], g['y']

    def _slot(self, shape, with_targets, slot=0):
        if self.skip_compute:
            return self._skip_state(shape, with_targets, slot)
        return self._capture(shape, with_targets, slot)

    def _replay_slot(self, g):
        if self.skip_compute:
            return self._run_skip_graphs(g)
        g['graph'].replay()
        return g['out']

    def _capture(self, shape, with_targets, slot=0):
        token = self._drop_stale_graphs()
        key = (shape, with_targets, slot, token)
        if key in self._graphs:
            return self._graphs[key]
        N, _, H, W = shape
        dev = self.device
        Xs = torch.zeros(shape, dtype=self.input_dtype, device=dev)
        ys = torch.full((N, 1, H, W), self.C, dtype=self.target_dtype, device=dev) if with_targets else None
        fn = (lambda: self._evaluate(Xs, ys)) if with_targets else (lambda: self._infer(Xs))
        saved = (self.cm.clone(), self.counts.clone(), self.exited_px.clone())
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(2):          # warm-up: lazy inits, plan cache, allocator
                fn()
        torch.cuda.current_stream(dev).wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = fn()
        # the warm-up / capture runs must not count
        self.cm.copy_(saved[0]); self.counts.copy_(saved[1]); self.exited_px.copy_(saved[2])
        # the folded plans the graph reads live outside its memory pool: keep them referenced next to it
        g = {'graph': graph, 'X': Xs, 'y': ys, 'out': out,
             'plans': (dict(getattr(self.net, '_plans', {})), dict(getattr(self.net, '_section_plans', {})))}
        self._graphs[key] = g
        return g

    def replay(self, shape, with_targets=True):
        return self._replay_slot(self._slot(tuple(shape), with_targets))

    def evaluate_pipelined(self, host_batches):
        """Throughput path for host-resident data: yields (exit int32 [N], scores f32 [E-1,N]) CPU
        tensors per batch, in order. `host_batches` yields (X, y) pinned CPU tensors of one shape.
        Two captured graphs with their own input buffers are used alternately: while graph k runs,
        the copy stream uploads batch k+1 into the other graph's buffers, and the small per-image
        results of batch k-1 are read back — H2D, compute and D2H overlap, nothing else changes.
        With skip_compute the "graph" of a slot is its chain of stage graphs (a 4-byte host read per gate)."""
        assert self.use_graph, "evaluate_pipelined needs use_graph=True"
        dev = self.device
        copy_stream = torch.cuda.Stream(device=dev)   # uploads
        back_stream = torch.cuda.Stream(device=dev)   # read-backs (must not queue in front of the next upload)
        main = torch.cuda.current_stream(dev)
        slots, pending = None, None
        ev_in = [torch.cuda.Event() for _ in range(2)]     # upload of slot k finished
        ev_free = [torch.cuda.Event() for _ in range(2)]   # graph of slot k finished reading its inputs
        ev_out = [torch.cuda.Event() for _ in range(2)]
        res_host = None

        def upload(k, X, y):
            nonlocal slots, res_host
            if slots is None:
                slots = [self._slot(tuple(X.shape), True, slot=i) for i in range(2)]
                res_host = [(torch.empty((X.shape[0],), dtype=torch.int32).pin_memory(),
                             torch.empty((max(self.E - 1, 1), X.shape[0]), dtype=torch.float32).pin_memory())
                            for _ in range(2)]
                for e in ev_free:
                    e.record(main)
            g = slots[k & 1]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_free[k & 1])
                g['X'].copy_(X, non_blocking=True)
                g['y'].copy_(y.view_as(g['y']), non_blocking=True)
                ev_in[k & 1].record(copy_stream)

        it = iter(host_batches)
        nxt = next(it, None)
        if nxt is not None:
            upload(0, *nxt)
        k = 0
        while nxt is not None:
            nxt = next(it, None)
            if nxt is not None:
                # slot (k+1)&1 was last read by batch k-1, whose results were handed out in the previous iteration
                upload(k + 1, *nxt)
            g = slots[k & 1]
            main.wait_event(ev_in[k & 1])
            self._replay_slot(g)              # skip mode blocks the host at every gate: the next upload is already queued
            ev_free[k & 1].record(main)
            with torch.cuda.stream(back_stream):
                back_stream.wait_event(ev_free[k & 1])
                res_host[k & 1][0].copy_(g['out']['exit'], non_blocking=True)
                res_host[k & 1][1].copy_(g['out']['scores'], non_blocking=True)
                ev_out[k & 1].record(back_stream)
            if pending is not None:
                ev_out[pending].synchronize()
                yield res_host[pending][0].clone(), res_host[pending][1].clone()
            pending = k & 1
            k += 1
        if pending is not None:
            ev_out[pending].synchronize()
            yield res_host[pending][0].clone(), res_host[pending][1].clone()

    # ------------------------------------------------------------------------------------------
    # compute-skipping + CUDA graphs: one graph per (stage i, active images n)
    # ------------------------------------------------------------------------------------------
    def _skip_state(self, shape, with_targets, slot=0):
        """Static buffers of the staged step for one input shape: the input of every backbone section
        (`xin[i]`, its first n rows hold the compacted still-active images), their original batch
        positions (`act[i]`), the per-image results, and a pinned host word per gate for the count."""
        token = self._drop_stale_graphs()
        key = ('skip', shape, with_targets, slot, token)
        st = self._graphs.get(key)
        if st is not None:
            return st
        N, _, H, W = shape
        dev, E = self.device, self.E
        X = torch.zeros(shape, dtype=self.input_dtype, device=dev)
        xin, Xc = [X], X
        for i in range(E - 1):                 # one eager pass: shapes of the section boundaries (+ plan warm-up)
            Xc = self.net.run_section(i, Xc)
            xin.append(torch.empty_like(Xc))
        st = {
            'X': X, 'xin': xin,
            'y': torch.full((N, 1, H, W), self.C, dtype=self.target_dtype, device=dev) if with_targets else None,
            'act': [torch.arange(N, dtype=torch.int64, device=dev)] +
                   [torch.zeros((N,), dtype=torch.int64, device=dev) for _ in range(E - 1)],
            'cnt_host': torch.zeros((E,), dtype=torch.int32).pin_memory(),
            'pool': torch.cuda.graph_pool_handle(), 'stages': {}, 'ev': torch.cuda.Event(),
            'out': {'exit': torch.full((N,), -1, dtype=torch.int32, device=dev),
                    'pred': torch.zeros((N, H, W), dtype=torch.uint8, device=dev),
                    'scores': torch.full((max(E - 1, 1), N), float('inf'), dtype=torch.float32, device=dev)},
        }
        self._graphs[key] = st
        return st

    def _skip_stage(self, st, i, n):
        """Stage i on the n still-active images: section, head, gate, decision; results scattered to the images'
        batch positions; survivors gathered to the front of the next section's input. Static shapes only."""
        net, E, out = self.net, self.E, st['out']
        N, H, W = out['pred'].shape
        last = i == E - 1
        if i == 0:
            out['exit'].fill_(-1)
            out['scores'].fill_(float('inf'))
        Xin, act = st['xin'][i][:n], st['act'][i][:n]
        Xc = net.run_section(i, Xin)
        low = net._plan(i).run(Xc)
        gated = not last and i >= self.skip
        res = self._gate(low, (H, W), want_score=gated)
        if not last and not gated:                  # an exit that is not allowed to answer: everybody moves on
            st['xin'][i + 1][:n].copy_(Xc)
            st['act'][i + 1][:n].copy_(act)
            return
        # one launch: decision, results of the leaving images to their batch positions, survivor list + count + positions
        al = torch.empty((n,), dtype=torch.int32, device=Xc.device)
        ac = torch.empty((1,), dtype=torch.int32, device=Xc.device)
        px = res.exited_px if gated else None
        with torch.cuda.device(Xc.device):
            ops.check(ops.lib().eeseg_exit_stage_commit(
                res.score.data_ptr() if gated else None, self.tau, 1, i, 1 if last else 0, act.data_ptr(),
                res.amax.data_ptr(), n, H * W, out['scores'][i].data_ptr() if gated else None, out['exit'].data_ptr(),
                out['pred'].data_ptr(), None if px is None else px.data_ptr(),
                None if px is None else self.exited_px[i:i + 1].data_ptr(), al.data_ptr(), ac.data_ptr(),
                None if last else st['act'][i + 1].data_ptr(), torch.cuda.current_stream(Xc.device).cuda_stream),
                "eeseg_exit_stage_commit")
        if not last:
            ops.compact_rows(self._dense_rows(Xc), al, ac, self._dense_rows(st['xin'][i + 1][:n]))   # survivors to the front
            st['cnt_host'][i:i + 1].copy_(ac, non_blocking=True)

    @staticmethod
    def _dense_rows(t):
        """[n, ...] view whose rows are dense in memory (channels_last NCHW tensors are NHWC underneath)."""
        if t.dim() == 4 and t.shape[0] and not t[0].is_contiguous() and t.is_contiguous(memory_format=torch.channels_last):
            return t.permute(0, 2, 3, 1)
        return t

    def _skip_final(self, st):
        out = st['out']
        cm = ops.confusion_hist(out['pred'], st['y'], self.C)            # [N, C+1, C]
        ex = out['exit'].long()
        self.cm.index_add_(0, ex, cm)
        self.cm[-1] += cm.sum(dim=0)
        self.counts.index_add_(0, ex, torch.ones_like(ex))
        self.counts[-1] += ex.numel()

    def _skip_graph(self, st, key, fn):
        g = st['stages'].get(key)
        if g is None:
            dev = self.device
            saved = (self.cm.clone(), self.counts.clone(), self.exited_px.clone())
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                for _ in range(2):      # warm-up on the real buffers: every stage is idempotent but for the accumulators
                    fn()
            torch.cuda.current_stream(dev).wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=st['pool']):
                fn()
            self.cm.copy_(saved[0]); self.counts.copy_(saved[1]); self.exited_px.copy_(saved[2])
            st['stages'][key] = g
            st.setdefault('plans', []).append((dict(getattr(self.net, '_plans', {})), dict(getattr(self.net, '_section_plans', {}))))
        return g

    def _run_skip_graphs(self, st):
        """Replay the stage graphs of one batch whose inputs are already in st['X'] / st['y']."""
        n = st['X'].shape[0]
        main = torch.cuda.current_stream(self.device)
        for i in range(self.E):
            self._skip_graph(st, (i, n), lambda: self._skip_stage(st, i, n)).replay()
            if i < self.E - 1 and i >= self.skip:
                st['ev'].record(main)
                st['ev'].synchronize()           # the one host read per gate: 4 bytes
                n = int(st['cnt_host'][i])
                if n == 0:
                    break
        if st['y'] is not None:
            self._skip_graph(st, 'final', lambda: self._skip_final(st)).replay()
        return st['out']

    @torch.no_grad()
    def infer(self, X):
        """X [N,3,H,W] on the device. Returns dict: 'exit' int32 [N] (0-based exit taken),
        'pred' uint8 [N,H,W] (argmax map of that exit), 'scores' f32 [E-1,N]. In graph mode the
        returned tensors are the graph's static outputs (overwritten by the next call)."""
        if self.use_graph and self.skip_compute:
            st = self._skip_state(tuple(X.shape), False)
            st['X'].copy_(X, non_blocking=True)
            return self._run_skip_graphs(st)
        if self.use_graph:
            g = self._capture(tuple(X.shape), False)
            g['X'].copy_(X, non_blocking=True)
            g['graph'].replay()
            return g['out']
        return self._infer(X)

    @torch.no_grad()
    def _infer(self, X, targets=None):
        if self.skip_compute:
            return self._infer_skipping(X)
        net = self.net
        N, _, H, W = X.shape
        dev = X.device
        E = self.E
        exit_idx = torch.full((N,), -1, dtype=torch.int32, device=dev)
        amax_all = torch.empty((E, N, H, W), dtype=torch.uint8, device=dev)
        scores = torch.full((max(E - 1, 1), N), float('inf'), dtype=torch.float32, device=dev)
        pool = self.metric != 'ent'
        # An early exit's head and gate (ASPP convolutions, up-sample + entropy + decision) do not feed the next backbone
        # section: they run on a side stream, forked when the section's output is ready and joined before the final
        # accumulation. Every conv launch is a persistent grid sized to the whole GPU, so two of them never share an SM —
        # but the tail of one (a 144-tile wave on 148 SMs, CTAs draining their epilogues) frees SMs that the other
        # stream's next launch fills, which a single in-order stream leaves idle. Captured into the CUDA graph as a
        # parallel branch; all early exits share the one side stream, so decisions stay ordered.
        # (overlap_gates without overlap_heads: only the gate runs aside, as in round 1.)
        main = torch.cuda.current_stream(dev)
        side = self._side_stream()
        keep = []          # tensors produced on one stream and read on the other stay referenced until the join
        forked = False
        Xc = X
        for i in range(E):
            Xc = net.run_section(i, Xc)
            gated = i < E - 1 and i >= self.skip
            head_aside = self.overlap_heads and self.overlap_gates and i < E - 1
            gate_aside = self.overlap_gates and i < E - 1
            if head_aside:
                side.wait_stream(main)
                keep.append(Xc)
                forked = True
            with torch.cuda.stream(side if head_aside else main):
                low = net._plan(i).run(Xc)
            if gate_aside and not head_aside:
                side.wait_stream(main)
                forked = True
            if gate_aside:
                keep.append(low)
            with torch.cuda.stream(side if gate_aside else main):
                res = ops.exit_gate(low, (H, W), layout='NHWC', n_classes=self.C, tau=self.tau,
                                    want_ent=pool and gated, want_amax=True, want_score=gated and not pool,
                                    amax_out=amax_all[i], score_out=scores[i] if gated and not pool else None)
                if gated:
                    if pool:
                        scores[i] = ops.entropy_pool_mean(res.ent, self.size, self.metric == 'min')
                    else:
                        self.exited_px[i] += res.exited_px.sum()
                    ops.gate_decide(scores[i], self.tau, i, exit_idx, want_active=False)
                if gate_aside:
                    keep.append(res)
        if forked:
            main.wait_stream(side)
        if targets is not None:
            # one kernel: final-exit assignment, histogram of the map each image took, accumulators, pred
            pred = torch.empty((N, H, W), dtype=torch.uint8, device=dev)
            tg = targets.reshape(N, -1)
            tg = (tg if tg.dtype in (torch.int64, torch.uint8) else tg.to(torch.int64)).contiguous()
            accumulate = ops.lib().eeseg_exit_accumulate_u8 if tg.dtype == torch.uint8 else ops.lib().eeseg_exit_accumulate
            with torch.cuda.device(dev):
                ops.check(accumulate(
                    amax_all.data_ptr(), tg.data_ptr(), exit_idx.data_ptr(), E, N, self.C, H * W,
                    self.cm.data_ptr(), self.counts.data_ptr(), pred.data_ptr(),
                    torch.cuda.current_stream(dev).cuda_stream), "eeseg_exit_accumulate")
            return {'exit': exit_idx, 'pred': pred, 'scores': scores}
        exit_idx = torch.where(exit_idx < 0, torch.full_like(exit_idx, E - 1), exit_idx)
        pred = amax_all[exit_idx.long(), torch.arange(N, device=dev)]     # the map of the exit each image took
        return {'exit': exit_idx, 'pred': pred, 'scores': scores}

    @torch.no_grad()
    def _infer_skipping(self, X):
        """Compute-skipping variant: after every gate the still-active images are compacted and only
        they run through the next section (one 4-byte D2H per exit for the active count)."""
        net = self.net
        N, _, H, W = X.shape
        dev = X.device
        exit_idx = torch.full((N,), -1, dtype=torch.int32, device=dev)
        pred = torch.empty((N, H, W), dtype=torch.uint8, device=dev)
        scores = torch.full((max(self.E - 1, 1), N), float('inf'), dtype=torch.float32, device=dev)
        active = None   # index tensor of images still in flight
        Xc = X
        for i in range(self.E):
            Xc = net.run_section(i, Xc)
            low = net._plan(i).run(Xc)
            last = i == self.E - 1
            res = self._gate(low, (H, W), want_score=not last and i >= self.skip)
            idx = active if active is not None else slice(None)
            if last:
                still = exit_idx[idx] < 0
                sel = still.view(-1, 1, 1)
                pred[idx] = torch.where(sel, res.amax, pred[idx])
                exit_idx[idx] = torch.where(still, torch.full_like(exit_idx[idx], i), exit_idx[idx])
                break
            if i >= self.skip:
                scores[i, idx] = res.score
                sub = exit_idx[idx].contiguous()
                before = sub < 0
                al, ac = ops.gate_decide(res.score, self.tau, i, sub, want_active=True)
                took = before & (sub == i)
                pred[idx] = torch.where(took.view(-1, 1, 1), res.amax, pred[idx])
                exit_idx[idx] = sub
                if res.exited_px is not None:
                    self.exited_px[i] += res.exited_px.sum()
                k = int(ac.item())              # one 4-byte D2H per exit
                if k == 0:
                    break
                if k < Xc.shape[0]:
                    keep = al[:k].long()
                    Xc = Xc[keep]
                    active = keep if active is None else active[keep]
        return {'exit': exit_idx, 'pred': pred, 'scores': scores}

    @torch.no_grad()
    def evaluate(self, X, y):
        """infer + integer confusion matrices of the exit taken: accumulates self.cm[e] for the exit
        each image left at and self.cm[-1] globally (the accumulators of eval_br_ent.py:39,61-69)."""
        if self.use_graph and self.skip_compute:
            st = self._skip_state(tuple(X.shape), True)
            st['X'].copy_(X, non_blocking=True)
            st['y'].copy_(y.view_as(st['y']), non_blocking=True)
            return self._run_skip_graphs(st)
        if self.use_graph:
            g = self._capture(tuple(X.shape), True)
            g['X'].copy_(X, non_blocking=True)
            g['y'].copy_(y.view_as(g['y']), non_blocking=True)
            g['graph'].replay()
            return g['out']
        return self._evaluate(X, y)

    @torch.no_grad()
    def _evaluate(self, X, y):
        if not self.skip_compute:
            return self._infer(X, targets=y)
        out = self._infer(X)
        cm = ops.confusion_hist(out['pred'], y, self.C)                  # [N, C+1, C]
        ex = out['exit'].long()
        self.cm.index_add_(0, ex, cm)
        self.cm[-1] += cm.sum(dim=0)
        self.counts.index_add_(0, ex, torch.ones_like(ex))
        self.counts[-1] += ex.numel()
        return out

    def all_reduce(self):
        """Sum the integer accumulators over ranks (NCCL): exact, order-independent."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.cm)
            dist.all_reduce(self.counts)
            dist.all_reduce(self.exited_px)

    def results(self):
        """Result dict with the reference's keys (eval_br_ent.py:72-84); mIoU from exact counts."""
        cm = self.cm.cpu()
        counts = self.counts.cpu()

        def miou(m):
            tp, fp, fn = ops.basics_from_cm(m)
            return float(((tp.double() / (tp + fp + fn).double()).sum() / self.C))
        res = {}
        for i in range(self.E - 1):
            res[f'b{i+1}_mIoU'] = miou(cm[i])
            res[f'b{i+1}_count'] = int(counts[i])
        res['mIoU_out'] = miou(cm[self.E - 1])
        res['count_out'] = int(counts[self.E - 1])
        res['mIoU_gl'] = miou(cm[-1])
        res['out_gl'] = int(counts[-1])
        res['t'] = self.tau
        res['pool'] = self.metric
        res['pool_size'] = self.size
        return res


class ThresholdSweep:
    """Single-pass threshold sweep: what re-running the reference's
    `br_evaluator` (eval_br_ent.py:38-84, CLI `-t`) once per tau computes, from ONE forward per batch.

    Per image the E-1 per-exit confidence scalars and the E confusion matrices (every exit's argmax
    against the target) are computed once; each tau is then resolved on those tiny tensors with the
    reference's rule (first early exit i >= skip whose score < tau, else the final exit). The integer
    accumulators `cm[T, E+1, C+1, C]` / `counts[T, E+1]` are exactly what T separate evaluations would
    give, and sum exactly across ranks (`all_reduce`)."""

    def __init__(self, net, n_classes, taus, metric='ent', size=1, skip=0):
        self.net = net
        self.C = n_classes
        self.E = net.n_branches + 1
        self.metric = metric.lower()
        assert self.metric in ('ent', 'max', 'min')
        self.size = size if self.metric != 'ent' else 1
        self.skip = skip
        dev = next(net.parameters()).device
        self.device = dev
        self.taus = torch.as_tensor(list(taus), dtype=torch.float32, device=dev)
        T = self.taus.numel()
        self.cm = torch.zeros((T, self.E + 1, n_classes + 1, n_classes), dtype=torch.int64, device=dev)
        self.counts = torch.zeros((T, self.E + 1), dtype=torch.int64, device=dev)

    @torch.no_grad()
    def update(self, X, y):
        H, W = X.shape[-2:]
        N = X.shape[0]
        pool = self.metric != 'ent'
        lows = self.net.forward_lowres(X)
        scores = torch.full((max(self.E - 1, 1), N), float('inf'), dtype=torch.float32, device=X.device)
        cms = []
        for i, lo in enumerate(lows):
            gated = i < self.E - 1 and i >= self.skip
            res = ops.exit_gate(lo, (H, W), layout='NHWC', n_classes=self.C, want_ent=pool and gated,
                                want_amax=True, want_score=gated and not pool)
            if gated:
                scores[i] = ops.entropy_pool_mean(res.ent, self.size, self.metric == 'min') if pool else res.score
            cms.append(ops.confusion_hist(res.amax, y, self.C))
        cms = torch.stack(cms)                                            # [E, N, C+1, C]
        # exit taken per (tau, image): first early exit whose score is below tau, else the last exit
        below = scores[: self.E - 1].unsqueeze(0) < self.taus.view(-1, 1, 1) if self.E > 1 else None  # [T, E-1, N]
        if below is not None and below.shape[1] > 0:
            first = torch.where(below.any(dim=1), below.float().argmax(dim=1), torch.full_like(below[:, 0], self.E - 1, dtype=torch.int64))
        else:
            first = torch.full((self.taus.numel(), N), self.E - 1, dtype=torch.int64, device=X.device)
        onehot = torch.nn.functional.one_hot(first, self.E).to(torch.int64)   # [T, N, E]
        # [T, E, C+1, C]; integer (CUDA has no int64 matmul): one masked sum per exit
        per_exit = torch.stack([(onehot[:, :, e, None, None] * cms[e].unsqueeze(0)).sum(dim=1) for e in range(self.E)], dim=1)
        self.cm[:, : self.E] += per_exit
        self.cm[:, self.E] += per_exit.sum(dim=1)
        self.counts[:, : self.E] += onehot.sum(dim=1)
        self.counts[:, self.E] += N
        return scores

    def all_reduce(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.cm)
            dist.all_reduce(self.counts)

    def results(self):
        """One result dict per tau with the reference's keys (eval_br_ent.py:72-84)."""
        cm, counts = self.cm.cpu(), self.counts.cpu()
        out = []
        for t, tau in enumerate(self.taus.cpu().tolist()):
            def miou(m):
                tp, fp, fn = ops.basics_from_cm(m)
                return float((tp.double() / (tp + fp + fn).double()).sum() / self.C)
            r = {}
            for i in range(self.E - 1):
                r[f'b{i+1}_mIoU'] = miou(cm[t, i])
                r[f'b{i+1}_count'] = int(counts[t, i])
            r['mIoU_out'] = miou(cm[t, self.E - 1])
            r['count_out'] = int(counts[t, self.E - 1])
            r['mIoU_gl'] = miou(cm[t, self.E])
            r['out_gl'] = int(counts[t, self.E])
            r['t'] = tau
            r['pool'] = self.metric
            r['pool_size'] = self.size
            out.append(r)
        return out
