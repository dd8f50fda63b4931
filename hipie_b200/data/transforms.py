"""Test-time resize of the predictor edge: detectron2's ResizeShortestEdge + ResizeTransform for uint8 images
(/root/reference/detectron2/data/transforms/augmentation_impl.py:120-195, transform.py:94-150: PIL bilinear)."""
import numpy as np


class ResizeShortestEdge:
    def __init__(self, short_edge_length, max_size=2 ** 31 - 1):
        if isinstance(short_edge_length, int):
            short_edge_length = (short_edge_length, short_edge_length)
        self.short_edge_length = tuple(short_edge_length)
        self.max_size = max_size

    @staticmethod
    def get_output_shape(oldh, oldw, short_edge_length, max_size):
        h, w = oldh, oldw
        size = short_edge_length * 1.0
        scale = size / min(h, w)
        if h < w:
            newh, neww = size, scale * w
        else:
            newh, neww = scale * h, size
        if max(newh, neww) > max_size:
            scale = max_size * 1.0 / max(newh, neww)
            newh = newh * scale
            neww = neww * scale
        neww = int(neww + 0.5)
        newh = int(newh + 0.5)
        return (newh, neww)

    def apply_image(self, img):
        """img (H, W, C) uint8 -> resized uint8 (PIL bilinear, as detectron2 does for uint8 inputs)."""
        from PIL import Image
        h, w = img.shape[:2]
        size = self.short_edge_length[0]          # test time: [min, min] range -> the single value
        if size == 0:
            return img
        newh, neww = self.get_output_shape(h, w, size, self.max_size)
        assert img.dtype == np.uint8, "the predictor takes uint8 images (cv2.imread layout)"
        pil = Image.fromarray(np.ascontiguousarray(img))
        return np.asarray(pil.resize((neww, newh), Image.BILINEAR))
