"""Module paths mirror the reference's ``core`` package (parse_model, yolo_decode_layer, yolo_nms, yolo_nms_layer, utils)."""
