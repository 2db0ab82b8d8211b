"""Two fused-classifier launches on 333 random ROIs (development tool: the command ncu wraps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import litepi_b200
from litepi_b200.classifier import _random_state_dict
clf = litepi_b200.B200Classifier(None, "shufflenetv2", num_classes=49, state_dict=_random_state_dict(49, 0, "shufflenetv2"), max_batch=512)
x = torch.randint(0, 255, (333, 64, 64, 3), dtype=torch.uint8, device=clf.device)
for _ in range(2):
    clf.classify_device(x)
torch.cuda.synchronize()
print("ok")
