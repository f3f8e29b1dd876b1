"""Row-sharded HMF (BASELINE config 4: HMF + WMRB `mw`, item tables sharded across up to 8 GPUs).

One process per GPU.  Every table row t lives on GPU t % G with its Adagrad accumulator; each rank
pools the rows it owns for the WHOLE global batch (the sm_100a kernels skip foreign rows), the
partial pooled vectors are reduce-scattered so that rank r holds the complete user / target-item
vectors of its own mb rows, scores and the loss are computed on those rows, and the gradients of
the pooled vectors are all-gathered back so that every owner updates its rows locally.
"""
import sys

import numpy as np
import torch

from .. import _lib
from .._lib import POOL_MEAN, OPT_ADAGRAD, call
from .hmf_model import LatentProductModel
from .exchange import RowShardExchange


class ShardedLatentProductModel(LatentProductModel):
    def __init__(self, *a, group=None, peer=None, **kw):
        """peer: True / False forces the NVLink peer-memory exchange on / off; None = on unless ARX_PEER=0."""
        self._peer_pref = peer
        self.ex = RowShardExchange(group)
        kw['shard'] = (self.ex.G, self.ex.r)
        super(ShardedLatentProductModel, self).__init__(*a, **kw)
        if self.loss_function != 'mw':
            raise NotImplementedError('the sharded path covers HMF with loss mw (config 4)')

    def step(self, session, user_input, item_input, neg_item_input=None, item_sampled=None,
             item_sampled_id2idx=None, forward_only=False, recommend=False, recommend_new=False, loss=None,
             run_op=None, run_meta=None, masks=None, sync=True):
        """user_input / item_input: the GLOBAL batch (G*mb ids, identical on every rank);
        rank r scores rows [r*mb, (r+1)*mb).  Returns the global mean loss."""
        if forward_only or recommend:
            return self._eval_or_recommend(user_input, item_input, recommend, recommend_new, loss)
        m, ex = self.att_emb, self.ex
        G, r, d = ex.G, ex.r, self.size
        dev = self.device
        m.add_input({}, user_input, item_input, item_sampled=item_sampled, item_sampled_id2idx=item_sampled_id2idx,
                    loss='mw')
        users_g = m.u_indices['input']
        items_g = m._ids(item_input)
        n_g = users_g.numel()
        mb = n_g // G
        S = m.sampled_ids.numel()
        pre = m._out_prefix()
        if self._peer_ok(mb, S, d):
            return self._step_peer(users_g, items_g, mb, S, masks, sync)
        # ---- forward: partial pooling of owned rows, one RS + one AR --------------------------
        # one packed exchange buffer per collective; row pitch 2d + 4 keeps every block 16-byte aligned so that the
        # lookups pool STRAIGHT into it (no staging copies): [ user vector | target-item vector | target bias, pad ]
        W = 2 * d + 4
        fwd = torch.empty((n_g, W), dtype=torch.float32, device=dev)
        sp = torch.empty((S, d + 4), dtype=torch.float32, device=dev)
        irng0 = m.sets[pre].attr_range()
        # the backward plans depend on the ids only: build them on side streams under the forward pass
        m.prefetch_plans({'user': [(m.sets['user'].attr_range(), users_g, POOL_MEAN)],
                          pre: [(irng0, m.sampled_ids, POOL_MEAN), (irng0, items_g, POOL_MEAN)]})
        (pu, _, urng), (pt, bt, irng), (ps, bs, _) = m.pool_many([
            ('user', users_g, POOL_MEAN, False, {'out': fwd[:, :d]}),
            (pre, items_g, POOL_MEAN, True, {'out': fwd[:, d:2 * d]}),
            (pre, m.sampled_ids, POOL_MEAN, True, {'out': sp[:, :d]})])
        fwd[:, 2 * d] = bt
        fwd[:, 2 * d + 1:] = 0
        sp[:, d] = bs
        sp[:, d + 1:] = 0
        loc = ex.reduce_scatter_rows(fwd)
        sp = ex.all_reduce(sp)
        U0 = loc[:, :d].contiguous()
        Pt = loc[:, d:2 * d].contiguous()
        btl = loc[:, 2 * d].contiguous()
        Ps = sp[:, :d].contiguous()
        bsl = sp[:, d].contiguous()
        users_l = users_g[r * mb:(r + 1) * mb].contiguous()
        keep = self.dropout
        scale = self._scale(n_g)[:mb]                                   # 1 / (G*mb): global batch mean
        mlp = self.nonlinear in ('relu', 'tanh')
        if (not mlp) and _lib.ce_supported(mb, S, d) and d % 4 == 0:
            # fused glue of the single-GPU step (arx_mw_prep / arx_mw_post): dropout (mask injected, or drawn in the
            # kernel by Philox) + tf32 rounding + transposes + target score in one launch, the two adjoints in another
            f32 = dict(dtype=torch.float32, device=dev)
            u = torch.empty((mb, d), **f32); U_r = torch.empty((mb, d), **f32); UT = torch.empty((d, mb), **f32)
            P_r = torch.empty((S, d), **f32); PT = torch.empty((d, S), **f32)
            tscore = torch.empty((mb,), **f32)
            inv_keep = 1.0 / keep
            dmask = drng = dmask_out = None
            if keep != 1.0:
                if masks:
                    dmask = masks[0]
                else:
                    if not hasattr(self, '_drop_rng'):
                        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) + 7919 * r        # a different stream per rank
                        self._drop_rng = torch.tensor([seed, 0], dtype=torch.int64, device=dev)
                    drng = self._drop_rng
                    dmask_out = torch.empty((mb, d), **f32)
            call('arx_mw_prep', U0.data_ptr(), _lib.ptr(dmask), inv_keep, _lib.ptr(drng), _lib.ptr(dmask_out), Pt.data_ptr(),
                 btl.data_ptr(), Ps.data_ptr(), mb, S, d, u.data_ptr(), U_r.data_ptr(), UT.data_ptr(), tscore.data_ptr(),
                 P_r.data_ptr(), PT.data_ptr())
            if drng is not None:
                dmask = dmask_out
            fused = m.fused_mw(u, Ps, bsl, tscore, scale, True, pos_rows=users_l, prepared=(U_r, P_r, UT, PT))
            if fused is None:
                raise RuntimeError('arx_mw_fwd / arx_mw_bwd rejected a shape ce_supported() accepted')
            bl, (dU, dPs, dbs, dts) = fused
            loss_sum = (bl.sum() / n_g).reshape(1)
            dU0 = torch.empty((mb, d), **f32)
            dPt = torch.empty((mb, d), **f32)
            call('arx_mw_post', dU.data_ptr(), dts.data_ptr(), Pt.data_ptr(), u.data_ptr(), _lib.ptr(dmask), inv_keep, mb, d,
                 dU0.data_ptr(), dPt.data_ptr(), _lib.ptr(drng))
        else:
            if mlp:
                # MLP tower (hmf_model.py:87-94) on this rank's rows; its dense parameters are replicated
                u0_leaf, u_graph = self._mlp_tower(U0, keep, masks)
                u = u_graph.detach().contiguous()
            else:
                u = m.dropout(U0, keep, masks[0] if masks else None)
                dmask = getattr(m, '_last_dropout_mask', None) if keep != 1.0 else None
            tscore = torch.empty((mb,), dtype=torch.float32, device=dev)
            call('arx_rowdot_fwd', u.data_ptr(), Pt.data_ptr(), btl.data_ptr(), mb, d, tscore.data_ptr())
            fused = m.fused_mw(u, Ps, bsl, tscore, scale, True, pos_rows=users_l)
            if fused is not None:                                           # scoring + WMRB + adjoints on the tensor cores
                bl, (dU, dPs, dbs, dts) = fused
            else:
                logits = torch.empty((mb, S), dtype=torch.float32, device=dev)
                _lib.gemm(u, Ps, logits, mb, S, d, 0, 1, bsl)
                bl = m.compute_loss(logits, tscore, 'mw', row_scale=scale, want_grad=True, pos_rows=users_l)
                # ---- backward: one AR + one AG, then local sparse Adagrad ----------------------------
                D, dts = logits, m._last_dtarget
                dU, dPs, dbs = self._scores_backward(D, u, Ps)
            loss_sum = (bl.sum() / n_g).reshape(1)
            dPt = torch.empty_like(Pt)
            call('arx_rowdot_bwd', u.data_ptr(), Pt.data_ptr(), dts.data_ptr(), mb, d, dU.data_ptr(), dPt.data_ptr())
            if mlp:
                for w in self.dense.values():
                    w.grad = None
                u_graph.backward(dU)
                dU0 = u0_leaf.grad
                # the ONE all-reduce of the dense parameters' gradients (every rank saw mb of the G*mb rows; the loss
                # is already scaled by 1 / (G*mb)): identical Adagrad updates on every replica
                names = [n for n, w in self.dense.items() if w.grad is not None]
                flat = torch.cat([self.dense[n].grad.reshape(-1) for n in names])
                ex.all_reduce(flat)
                o = 0
                for n in names:
                    g = self.dense[n].grad
                    g.copy_(flat[o:o + g.numel()].view_as(g))
                    o += g.numel()
            elif keep != 1.0:
                dU0 = torch.empty_like(dU)
                call('arx_scale_mask', dU.data_ptr(), dmask.data_ptr(), 1.0 / keep, dU.numel(), dU0.data_ptr())
            else:
                dU0 = dU
        dsp_in = torch.empty((S, d + 4), dtype=torch.float32, device=dev)
        dsp_in[:, :d] = dPs
        dsp_in[:, d] = dbs
        dsp_in[:, d + 1:] = 0
        dsp = ex.all_reduce(dsp_in)
        mine = torch.empty((mb, W), dtype=torch.float32, device=dev)
        mine[:, :d] = dU0
        mine[:, d:2 * d] = dPt
        mine[:, 2 * d] = dts
        mine[:, 2 * d + 1:] = 0
        back = ex.all_gather_rows(mine)
        # the gathered gradient rows are used in place (row-strided views: the kernels take a row pitch)
        m.push_grad('user', urng, users_g, POOL_MEAN, back[:, :d])
        rng = m.sets[pre].attr_range()
        m.push_grad(pre, rng, m.sampled_ids, POOL_MEAN, dsp[:, :d].contiguous(), dsp[:, d].contiguous())
        m.push_grad(pre, rng, items_g, POOL_MEAN, back[:, d:2 * d].contiguous(), back[:, 2 * d].contiguous())
        lr = self.learning_rate.eval()
        m.apply_gradients(lr, OPT_ADAGRAD)
        if mlp:
            for name, w in self.dense.items():
                if w.grad is not None:
                    call('arx_dense_update', w.data.data_ptr(), self.dense_acc[name].data_ptr(),
                         w.grad.contiguous().data_ptr(), w.numel(), float(lr), None, OPT_ADAGRAD)
        self.global_step.assign(self.global_step.eval() + 1)
        if not sync:
            return loss_sum
        ex.all_reduce(loss_sum)
        return float(loss_sum.item())

    # ---- evaluation / recommendation on row-sharded tables (hmf_model.py:193-211) ------------------------------
    def _catalog_full(self):
        """Pooled catalog [V, d] + bias [V]: every rank pools the rows it owns for ALL items, one all-reduce completes
        them; cached until the next training step (the evaluation loop calls this once per pass)."""
        key = self.global_step.eval()
        cache = getattr(self, '_cat_cache', None)
        if cache is None or cache[0] != key:
            P, beta, _ = self.att_emb.pool_catalog('full', 1)
            self.ex.all_reduce(P)
            self.ex.all_reduce(beta)
            self._cat_cache = (key, P, beta)
        return self._cat_cache[1], self._cat_cache[2]

    def _eval_or_recommend(self, user_input, item_input, recommend, recommend_new, loss=None):
        """user_input / item_input: the GLOBAL batch (identical on every rank, a multiple of G rows).  The user vectors
        are completed by an all-reduce of the per-rank partial pools, rank r scores rows [r*mb, (r+1)*mb) against the
        full catalog, and the per-row results are gathered: recommend returns int[G*mb, top_N] on every rank, evaluation
        the global mean of the reference's loss_eval ('warp' over the whole catalog, no positives masked: see
        LatentProductModel.step)."""
        if recommend and recommend_new:
            raise AttributeError("'LatentProductModel' object has no attribute 'indices_test'")       # hmf_model.py:198
        m, ex = self.att_emb, self.ex
        G, r, d = ex.G, ex.r, self.size
        m.add_input({}, user_input, item_input, forward_only=not recommend, recommend=recommend, loss='mw')
        users_g = m.u_indices['input']
        n_g = users_g.numel()
        assert n_g % G == 0, 'the global batch must be a multiple of the number of ranks'
        mb = n_g // G
        pu, _, _ = m.pool('user', users_g, POOL_MEAN, False)               # partial sums over the rows this rank owns
        ex.all_reduce(pu)
        u = pu[r * mb:(r + 1) * mb].contiguous()                           # keep_prob 1.0 (:167-170, :78)
        if self.nonlinear in ('relu', 'tanh'):
            with torch.no_grad():
                u = self._mlp_tower(u, 1.0, None)[1].contiguous()
        P, beta = self._catalog_full()
        V = P.shape[0]
        logits = torch.empty((mb, V), dtype=torch.float32, device=self.device)
        _lib.gemm(u, P, logits, mb, V, d, 0, 1, beta)                      # :118
        if recommend:
            idx = torch.empty((mb, self.top_N_items), dtype=torch.int32, device=self.device)
            call('arx_topk_rows', logits.data_ptr(), mb, V, logits.stride(0), self.top_N_items, idx.data_ptr(), None)
            return ex.all_gather_rows(idx).cpu().numpy()                   # :154, :200
        items_g = m._ids(item_input)
        targets = m.item2logit_dev[items_g[r * mb:(r + 1) * mb].long()].contiguous()
        # as LatentProductModel.step: the model's own sampled loss evaluates as the full-catalog 'warp' with NO positives
        # masked (the reference's quirk, hmf_model.py:209-211); an explicit loss (the runner passes 'warp') masks the
        # evaluation positives of each row's user
        eff = loss if loss is not None else self.loss_function
        unmasked = eff == 'mw'
        if eff == 'mw':
            eff = 'warp'
        users_l = users_g[r * mb:(r + 1) * mb].contiguous()
        bl = m.compute_loss(logits, targets, eff, loss_func=self.loss_func, exp_p=self.loss_exp_p, want_grad=False,
                            forward_only=True, unmasked=unmasked, pos_rows=users_l)
        total = (bl.sum() / n_g).reshape(1)
        ex.all_reduce(total)
        return float(total.item())

    # ---- the same step over NVLink peer memory (hmf/exchange.py::PeerExchange) ---------------------------------
    def _peer_ok(self, mb, S, d):
        """Peer-memory exchange: NCCL backend (one GPU per rank on one node), the fused glue's shapes, not switched off
        (ARX_PEER=0 keeps the NCCL collectives for A/B runs)."""
        if getattr(self, 'px', None) is not None:
            return self.px.mb == mb and self.px.S == S
        if getattr(self, '_peer_tried', False):
            return False
        self._peer_tried = True
        import os
        want = self._peer_pref if self._peer_pref is not None else os.environ.get('ARX_PEER', '1') == '1'
        if not self.ex.nccl or not want or self.att_emb.dim not in (128, 256) or self.nonlinear in ('relu', 'tanh'):
            return False
        if not (_lib.ce_supported(mb, S, d) and d % 4 == 0):
            return False
        from .exchange import PeerExchange, PeerUnavailable
        try:
            self.px = PeerExchange(self.ex.group, self.device, mb, S, d)
        except PeerUnavailable as e:             # raised on every rank alike: all of them keep the NCCL collectives
            if self._peer_pref:
                raise
            if self.ex.r == 0:
                print('[arecsys_b200] peer-memory exchange unavailable, using NCCL collectives: %s' % e, file=sys.stderr)
            self.px = None
            return False
        # no NCCL kernel is left in the captured step: use the single-GPU scheduling (plan builds on low-priority
        # streams, everything on the dependent chain — the barrier kernels above all — on high-priority ones)
        self.att_emb.use_priorities = os.environ.get('ARX_PEER_PRIO', '1') == '1'
        self.att_emb._side_streams = {}
        return True

    def _step_peer(self, users_g, items_g, mb, S, masks, sync):
        m, px = self.att_emb, self.px
        G, r, d = px.G, px.r, self.size
        dev = self.device
        n_g = users_g.numel()
        pre = m._out_prefix()
        f32 = dict(dtype=torch.float32, device=dev)
        px.clear_pool_gradients()                                   # before this step's first barrier
        # nothing below depends on the lookups: the positives bit matrix and the zeroed outputs of the split accumulation
        # are prepared on a side stream UNDER the lookups (forked here, before they are enqueued); they were 25 us of
        # memsets + mask build on the dependent chain
        users_l = users_g[r * mb:(r + 1) * mb].contiguous()
        main = torch.cuda.current_stream()
        side = m.side_stream(3)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            dU_z = torch.zeros((mb, d), **f32)
            dPs_z = torch.zeros((S, d), **f32)
            zb = torch.zeros((S + mb,), **f32)                      # dbs | dts
            mwmask = m.mw_mask(mb, S, pos_rows=users_l)
            ev_side = torch.cuda.Event()
            ev_side.record(side)
        irng0 = m.sets[pre].attr_range()
        m.prefetch_plans({'user': [(m.sets['user'].attr_range(), users_g, POOL_MEAN)],
                          pre: [(irng0, m.sampled_ids, POOL_MEAN), (irng0, items_g, POOL_MEAN)]})
        # ONE lookup launch: partial sums of the rows this rank owns, for ALL G*mb bags, added straight into the owners'
        # locU / locP / locb (lookup + reduce-scatter) and, for the pool, into every rank's spP / spb (lookup + all-reduce)
        (_, _, urng), (_, _, irng), _ = m.pool_many([
            ('user', users_g, POOL_MEAN, False, {'push': px.push_desc('user')}),
            (pre, items_g, POOL_MEAN, True, {'push': px.push_desc('item')}),
            (pre, m.sampled_ids, POOL_MEAN, True, {'push': px.push_desc('pool')})])
        px.barrier()                                                # B1: every rank's pushes have landed
        U0, Pt, btl, Ps, bsl = px.locU, px.locP, px.locb, px.spP, px.spb
        keep = self.dropout
        scale = self._scale(n_g)[:mb]                               # 1 / (G*mb): global batch mean
        u = torch.empty((mb, d), **f32); U_r = torch.empty((mb, d), **f32); UT = torch.empty((d, mb), **f32)
        P_r = torch.empty((S, d), **f32); PT = torch.empty((d, S), **f32)
        tscore = torch.empty((mb,), **f32)
        inv_keep = 1.0 / keep
        dmask = drng = dmask_out = None
        if keep != 1.0:
            if masks:
                dmask = masks[0]
            else:
                if not hasattr(self, '_drop_rng'):
                    seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) + 7919 * r            # a different stream per rank
                    self._drop_rng = torch.tensor([seed, 0], dtype=torch.int64, device=dev)
                drng = self._drop_rng
                dmask_out = torch.empty((mb, d), **f32)
        call('arx_mw_prep', U0.data_ptr(), _lib.ptr(dmask), inv_keep, _lib.ptr(drng), _lib.ptr(dmask_out), Pt.data_ptr(),
             btl.data_ptr(), Ps.data_ptr(), mb, S, d, u.data_ptr(), U_r.data_ptr(), UT.data_ptr(), tscore.data_ptr(),
             P_r.data_ptr(), PT.data_ptr())
        if drng is not None:
            dmask = dmask_out
        # fused_mw keeps no reference to Ps / bsl beyond its launches; they are read in place from the receive blocks
        main.wait_event(ev_side)
        fused = m.fused_mw(u, Ps, bsl, tscore, scale, True, pos_rows=users_l, prepared=(U_r, P_r, UT, PT), mask=mwmask,
                           outputs=(dU_z, dPs_z, zb[:S], zb[S:]))
        if fused is None:
            raise RuntimeError('arx_mw_fwd / arx_mw_bwd rejected a shape ce_supported() accepted')
        bl, (dU, dPs, dbs, dts) = fused
        side.wait_stream(main)                                      # the scalar loss is off the dependent chain
        with torch.cuda.stream(side):
            loss_sum = (bl.sum() / n_g).reshape(1)
            ev_loss = torch.cuda.Event()
            ev_loss.record(side)
        dU0 = torch.empty((mb, d), **f32)
        dPt = torch.empty((mb, d), **f32)
        call('arx_mw_post', dU.data_ptr(), dts.data_ptr(), Pt.data_ptr(), u.data_ptr(), _lib.ptr(dmask), inv_keep, mb, d,
             dU0.data_ptr(), dPt.data_ptr(), _lib.ptr(drng))
        px.push_gradients(dU0, dPt, dts, dPs.contiguous(), dbs.contiguous())      # backward exchange: one launch
        px.clear_forward_blocks()                                   # consumed: cleared before this step's second barrier
        px.barrier()                                                # B2
        # the receive blocks ARE the gradient arenas of the scatter-Adagrad kernels (no gather, no concatenation)
        rng = m.sets[pre].attr_range()
        m.push_grad('user', urng, users_g, POOL_MEAN, px.ug)
        m.push_grad(pre, rng, m.sampled_ids, POOL_MEAN, px.ig[:S], px.igb[:S])
        m.push_grad(pre, rng, items_g, POOL_MEAN, px.ig[S:], px.igb[S:])
        m.apply_gradients(self.learning_rate.eval(), OPT_ADAGRAD)
        main.wait_event(ev_loss)
        self.global_step.assign(self.global_step.eval() + 1)
        if not sync:
            return loss_sum
        self.ex.all_reduce(loss_sum)
        return float(loss_sum.item())

    def replay_step(self, users, items, sync=True):
        """Captured sharded step (the NCCL exchanges are part of the graph); users / items are the
        GLOBAL batch.  With sync the global mean loss is all-reduced and returned as a float; sync='lag' returns
        the previous step's (see LatentProductModel.replay_step): no host wait on the step just launched."""
        if sync is False:
            return super(ShardedLatentProductModel, self).replay_step(users, items, sync=False)
        return super(ShardedLatentProductModel, self).replay_step(users, items, sync=sync)

    def _loss_for_host(self):
        # the rank-local partial of the global mean -> global mean (one 4-byte all-reduce, stream-ordered)
        return self.ex.all_reduce(self._g_loss.clone())
