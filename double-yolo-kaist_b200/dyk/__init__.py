"""Runtime of the B200 dual-stream YOLO hot path (native binding, cfg zoo, execution plan)."""
