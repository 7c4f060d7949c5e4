#!/usr/bin/env python
"""Golden vectors from the UNMODIFIED /root/reference/BS_brain.py executed in THIS container on tests/keras_shim (a
stand-in for Keras 2.2.4 / TF 1.14 restating the handful of primitives the reference calls; tests/keras_shim/README.md).

    python tests/golden/make_refshim_golden.py [reference [out]]   # writes tests/golden/refshim_n4_b{1,64}.npz

Same driver, same file format and same consumer tests (tests/test_tf1_golden.py) as make_tf1_golden.py; the only
difference is what sits underneath `import keras`.  What executes is the reference's own model code: GNNLayer.build/call
(BS_brain.py:24-51), AggLayer.call (:69-76), BS._create_model (:108-214), train_dnn/predict/update_target_model
(:218-239).  /root/reference does not travel to the GPU box, hence committed vectors."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "keras_shim"))
sys.path.insert(0, HERE)
import make_tf1_golden as G     # noqa: E402

if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = sys.argv[2] if len(sys.argv) > 2 else HERE
    G.generate(ref, out, [1, 64], prefix="refshim", require_pinned_stack=False)
