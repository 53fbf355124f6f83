"""Visualisation entry points of the reference (myolo/visualize.py) -- plotting only, out of the
hot-path scope; they are no-ops unless matplotlib is installed."""
from mrcnn.visualize import display_images, display_instances, display_top_masks, draw_boxes, random_colors  # noqa: F401
