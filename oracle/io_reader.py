"""TEST INFRASTRUCTURE ONLY - literal restatement of the reference's feature record decoding, used to check
gst_visdial_b200/io/features.py.  Follows utils/image_features_reader.py:110-141 (the not-in-memory branch) and
utils/data_utils.py:73-117 (mask_prob = 0) statement by statement, with explicit loops."""
import base64
import copy

import numpy as np


def read_record(item):
    image_h = int(item['image_h'])
    image_w = int(item['image_w'])
    num_boxes = int(item['num_boxes'])
    features = np.frombuffer(base64.b64decode(item["features"]), dtype=np.float32).reshape(num_boxes, 2048)
    boxes = np.frombuffer(base64.b64decode(item['boxes']), dtype=np.float32).reshape(num_boxes, 4)
    g_feat = np.sum(features, axis=0) / num_boxes
    num_boxes = num_boxes + 1
    features = np.concatenate([np.expand_dims(g_feat, axis=0), features], axis=0)
    image_location = np.zeros((boxes.shape[0], 5), dtype=np.float32)
    image_location[:, :4] = boxes
    for i in range(boxes.shape[0]):
        image_location[i, 4] = (image_location[i, 3] - image_location[i, 1]) * (image_location[i, 2] - image_location[i, 0]) / (float(image_w) * float(image_h))
    _ = copy.deepcopy(image_location)
    for i in range(boxes.shape[0]):
        image_location[i, 0] = image_location[i, 0] / float(image_w)
        image_location[i, 1] = image_location[i, 1] / float(image_h)
        image_location[i, 2] = image_location[i, 2] / float(image_w)
        image_location[i, 3] = image_location[i, 3] / float(image_h)
    g_location = np.array([0, 0, 1, 1, 1])
    image_location = np.concatenate([np.expand_dims(g_location, axis=0), image_location], axis=0)
    return features, num_boxes, image_location


def encode_image_input(features, num_boxes, boxes, max_regions=37):
    num_boxes = min(int(num_boxes), max_regions)
    mix_boxes_pad = np.zeros((max_regions, boxes.shape[-1]))
    mix_features_pad = np.zeros((max_regions, features.shape[-1]))
    mix_boxes_pad[:num_boxes] = boxes[:num_boxes]
    mix_features_pad[:num_boxes] = features[:num_boxes]
    image_mask = [1] * (int(num_boxes))
    while len(image_mask) < max_regions:
        image_mask.append(0)
    return mix_features_pad.astype(np.float32), mix_boxes_pad.astype(np.float32), np.array(image_mask, dtype=np.float32)
