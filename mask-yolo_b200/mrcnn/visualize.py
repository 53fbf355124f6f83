"""mrcnn.visualize shim: plotting helpers degrade to no-ops without matplotlib."""


def _noop(*args, **kwargs):
    return None


display_images = display_instances = display_top_masks = draw_boxes = _noop


def random_colors(n, bright=True):
    import colorsys
    return [colorsys.hsv_to_rgb(i / max(n, 1), 1, 1.0 if bright else 0.7) for i in range(n)]
