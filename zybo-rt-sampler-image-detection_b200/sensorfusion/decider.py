"""Host mirror of the reference's PC/sensorfusion/decider.py for the parts that touch the
beamformer's output: the entropy confidence of a heat map (16-24, on the GPU) and the
box -> steering-angle mapping (70-81, host arithmetic).  Image compositing (create_image,
get_lightlevel) is UI and out of scope."""
import numpy as np

from lib import _native


class sensorfusiondecider:
    def __init__(self, display_size=(640, 360), MAX_ANGLE=30, ASPECT_RATIO=16 / 9):
        self.display_size = display_size
        self.image_confidence_threshold = 0.5
        self.MAX_X = MAX_ANGLE
        self.MAX_Y = MAX_ANGLE / ASPECT_RATIO

    def get_entropy(self, heatmap):
        """decider.py:16-24 for an 8-bit heat map (any shape); computed on the device."""
        import torch
        hm = np.ascontiguousarray(heatmap)
        if hm.dtype != np.uint8:
            raise TypeError("get_entropy: the device path takes the uint8 heat map the reference passes in")
        d = torch.from_numpy(hm).cuda()
        conf = torch.empty(1, dtype=torch.float64, device="cuda")
        _native.check(_native.lib().bf_entropy_dev(d.data_ptr(), 1, hm.size, conf.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream))
        return float(conf.cpu()[0])

    def focus_beam(self, callback, box):
        """decider.py:70-81."""
        x1, y1, x2, y2, conf = box
        if conf < self.image_confidence_threshold:
            return -1, -1
        x_mid = (x1 + x2) / 2
        y_mid = (y1 + y2) / 2
        horizontal = (x_mid / self.display_size[0]) * self.MAX_X * 2 - self.MAX_X
        vertical = (y_mid / self.display_size[1]) * self.MAX_Y * 2 - self.MAX_Y
        callback(horizontal, vertical)
        return 0
