"""B200-native implementation of the img2sgf diagram-recognition hot path.

Public entry points (mirroring /root/reference/img2sgf.py Part 2, see api.py):
    edge_map, find_circles, find_lines, find_all_lines, cluster, validate_grid, classify_stones,
    process_image; batch.Engine / batch.BatchRunner for batches and multi-GPU sharding.
"""
__version__ = "0.1.0"
