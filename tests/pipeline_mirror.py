"""Oracle-side mirror of the fb_update pipeline (flame_ros_b200/csrc/flame_update.cuh): the same
stage order and bookkeeping rules, with every numeric stage done by the CPU oracle.  Test
infrastructure only.  The triangulation is shared (the product's host-side fb_delaunay) because the
parity contract feeds both sides the same edge list; the triangulator itself is
checked against Qhull in tests/test_delaunay.py."""
import numpy as np


class MirrorFlame:
    def __init__(self, O, capi, W, H, K, n_slots, max_features, max_vertices, up, ep=None):
        self.O, self.capi, self.W, self.H, self.K = O, capi, W, H, np.asarray(K, np.float32)
        self.n_slots, self.maxF, self.maxV, self.up = n_slots, max_features, max_vertices, up
        self.ep = ep or O.EpiParams.default()
        self.imgs = np.zeros((n_slots, H, W), np.uint8)
        self.poses = np.zeros((n_slots, 7), np.float32)
        self.poses[:, 3] = 1.0
        F = max_features
        self.u_ref = np.zeros((F, 2), np.float32)
        self.ref_slot = np.zeros(F, np.int32)
        self.mu = np.zeros(F, np.float32)
        self.var = np.zeros(F, np.float32)
        self.dropouts = np.zeros(F, np.int32)
        self.alive = np.zeros(F, np.int32)
        self.valid = np.zeros(F, np.int32)
        self.u_cur = np.full((F, 2), np.nan, np.float32)
        self.pf_img_id = [-1] * (n_slots - 1)
        self.pf_next, self.have_pf = 0, False
        self.vert_feat, self.edges, self.tris, self.state, self.idmap, self.pos = None, None, None, None, None, None

    def _new_poseframe(self, img_id):
        up, cur, slot = self.up, self.n_slots - 1, self.pf_next
        if self.pf_img_id[slot] >= 0:
            self.alive[(self.alive == 1) & (self.ref_slot == slot)] = 0
        self.imgs[slot] = self.imgs[cur]
        self.poses[slot] = self.poses[cur]
        self.pf_img_id[slot] = img_id
        self.pf_next = (slot + 1) % (self.n_slots - 1)
        self.have_pf = True
        win = up.detection_win_size
        cx, cy = self.W // win, self.H // win
        occ = np.zeros(cx * cy, np.uint8)
        for f in np.nonzero(self.valid)[0]:
            i, j = int(np.floor(self.u_cur[f, 0])) // win, int(np.floor(self.u_cur[f, 1])) // win
            if 0 <= i < cx and 0 <= j < cy:
                occ[j * cx + i] = 1
        mag = self.O.gradient_mag(self.imgs[cur])
        _, det_xy, det_ok = self.O.detect_features(mag, win, up.detection_border, up.min_grad_mag, occ)
        free = np.nonzero(self.alive == 0)[0]
        dets = np.nonzero(det_ok)[0]
        for k in range(min(len(free), len(dets))):
            f, p = free[k], det_xy[dets[k]]
            m = np.float32(up.idepth_init)
            if up.init_with_prediction and self.idmap is not None:
                q = self.idmap[int(p[1]), int(p[0])]
                if np.isfinite(q) and q > 0:
                    m = q
            self.u_ref[f], self.ref_slot[f], self.mu[f], self.var[f] = p, slot, m, up.idepth_var_init
            self.dropouts[f], self.alive[f] = 0, 1

    def update(self, time, img_id, pose, gray, is_poseframe):
        O, up, cur = self.O, self.up, self.n_slots - 1
        self.imgs[cur] = gray
        self.poses[cur] = pose
        if not self.have_pf:
            self.valid[:] = 0
            self._new_poseframe(img_id)
            return False
        O.idepth_update(self.imgs, self.poses, self.K, cur, self.ref_slot, self.u_ref, self.mu, self.var,
                        self.dropouts, self.alive, self.ep)
        u_cur, mu_cur, var_cur, valid = O.project_features(self.W, self.H, self.poses, self.K, cur, self.ref_slot,
                                                           self.u_ref, self.mu, self.var, self.alive)
        self.alive[(self.alive == 1) & (valid == 0)] = 0
        self.valid, self.u_cur, self.mu_cur, self.var_cur = valid, u_cur, mu_cur, var_cur
        vfeat = [f for f in range(self.maxF) if valid[f] and var_cur[f] < up.idepth_var_max_graph][:self.maxV]
        V = len(vfeat)
        updated, have_tri = False, False
        if V >= 3:
            pos = np.ascontiguousarray(u_cur[vfeat], np.float32)
            try:
                tris, edges = self.capi.delaunay(pos)
                have_tri = True
            except self.capi.FlameError:
                have_tri = False
        if have_tri:
            E = len(edges)
            z = mu_cur[vfeat].astype(np.float32)
            wt = (np.float32(1.0) / var_cur[vfeat]).astype(np.float32) if up.adaptive_data_weights else np.ones(V, np.float32)
            st = {k: np.zeros(V, np.float32) for k in ("x", "w1", "w2", "xb", "w1b", "w2b")}
            st.update({k: np.zeros(E, np.float32) for k in ("q1", "q2", "q3")})
            f2v, eold = {}, {}
            if self.vert_feat is not None:
                f2v = {f: k for k, f in enumerate(self.vert_feat)}
                eold = {(self.vert_feat[a], self.vert_feat[b]): e for e, (a, b) in enumerate(self.edges)}
            for k, f in enumerate(vfeat):
                o = f2v.get(f, -1)
                if o >= 0:
                    for key in ("x", "w1", "w2", "xb", "w1b", "w2b"):
                        st[key][k] = self.state[key][o]
                else:
                    x0 = z[k]
                    if up.init_with_prediction and self.idmap is not None:
                        px, py = int(np.rint(pos[k, 0])), int(np.rint(pos[k, 1]))
                        if 0 <= px < self.W and 0 <= py < self.H:
                            p = self.idmap[py, px]
                            if np.isfinite(p) and p > 0:
                                x0 = p
                    st["x"][k] = st["xb"][k] = x0
            for e, (a, b) in enumerate(edges):
                o = eold.get((vfeat[a], vfeat[b]), -1)
                if o >= 0:
                    for key in ("q1", "q2", "q3"):
                        st[key][e] = self.state[key][o]
            d = pos[edges[:, 0]] - pos[edges[:, 1]]
            alpha = (np.float32(1.0) / np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])).astype(np.float32)
            beta = np.ones(E, np.float32)
            if up.do_nltgv2 and up.iters > 0:
                O.nltgv2_solve(pos, edges, alpha, beta, z, wt, st, up.rparams, up.iters)
            self.idmap = O.rasterize_idepth(self.W, self.H, pos, st["x"], tris, None)
            self.vert_feat, self.edges, self.tris, self.state, self.pos = vfeat, edges, tris, st, pos
            updated = True
        if is_poseframe:
            self._new_poseframe(img_id)
        return updated
