"""Per-frame hand tracking: the recurrence of the reference's HandTrackModel.forward
(network/models/track_network.py:139-226, the branch without IKNet / pose optimiser).

    frame t:   jittered_kp = last_kp + mean(hand_points_t)          (:163, "this trick is important for fast motion")
               ret = HandTrackNet(hand_points_t, jittered_kp)        (:215, track_flag=True, eval mode)
               last_kp = ret['pred_kp'] - mean(hand_points_t)        (:217)

Frame t+1 depends on frame t, the batch is one sequence (B = 1): the loop is launch-latency bound.  With the GPU Kabsch
there is no host synchronisation inside a frame, so the whole frame -- recurrence update included -- is captured ONCE in a
CUDA graph and replayed per frame; the state (last_kp) lives on the device between replays.
"""
import torch

from .hand_network import visibility_mask


class HandTracker:
    def __init__(self, handnet, palm_template, graph=True, visibility=False):
        """handnet: hotrack_b200.hand_network.HandTrackNet (or anything with its forward contract);
        palm_template [1|B, 6, 3]: canonical palm keypoints (the reference takes them from the MANO layer at zero pose,
        track_network.py:151-153)."""
        self.net = handnet.eval()
        self.palm = palm_template
        self.use_graph = graph
        self.visibility = visibility
        self.last_kp = None           # [B,21,3], relative to the cloud centroid
        self._graph = self._pts = self._out = self._vis = None

    def reset(self, init_kp, hand_points):
        """Start a sequence from the first frame's (jittered) keypoints, as the data loader provides them."""
        self.last_kp = (init_kp.float() - hand_points.float().mean(dim=-2, keepdim=True)).clone()

    @torch.no_grad()
    def _frame(self, hand_points):
        centre = hand_points.mean(dim=-2, keepdim=True)
        data = {'pred_palm_template': self.palm, 'hand_points': hand_points, 'jittered_hand_kp': self.last_kp + centre}
        ret = self.net(data, {'track_flag': True, 'test_flag': True, 'IKNet_flag': False})
        pred = ret['pred_kp']
        self.last_kp.copy_(pred - centre)
        vis = visibility_mask(pred, hand_points) if self.visibility else None
        return pred, vis

    @torch.no_grad()
    def step(self, hand_points):
        """hand_points [B,N,3] (CUDA, fixed shape across the sequence) -> pred_kp [B,21,3] (and the visibility mask
        [B,21] when enabled).  The returned tensors are overwritten by the next step when a graph is replayed."""
        if self.last_kp is None:
            raise RuntimeError("call reset(init_kp, hand_points) before the first step")
        if not self.use_graph:
            out = self._frame(hand_points.float())
            return out if self.visibility else out[0]
        if self._graph is None:
            self._pts = hand_points.float().clone()
            keep = self.last_kp.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up (lazy initialisations), state restored afterwards
                for _ in range(3):
                    self._frame(self._pts)
                self.last_kp.copy_(keep)
            torch.cuda.current_stream().wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._out, self._vis = self._frame(self._pts)
            self.last_kp.copy_(keep)  # capture does not execute, but keep the invariant explicit
        self._pts.copy_(hand_points, non_blocking=True)
        self._graph.replay()
        return (self._out, self._vis) if self.visibility else self._out

    def track(self, frames, init_kp):
        """frames: iterable of hand_points [B,N,3] -> list of pred_kp clones, one per frame."""
        out = []
        for i, pts in enumerate(frames):
            if i == 0:
                self.reset(init_kp, pts)
            r = self.step(pts)
            out.append((r[0] if self.visibility else r).clone())
        return out
