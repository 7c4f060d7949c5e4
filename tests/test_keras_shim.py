"""The Keras/TF stand-in (tests/keras_shim) against independent implementations of the same published rules: the
recordings tests/golden/refshim_*.npz are only as good as the primitives underneath the reference's model code, so each
primitive the reference calls is pinned here to torch's own implementation (third-party code) or to a closed form.
Runs in a fresh interpreter per check so that the fake `keras` / `tensorflow` modules never enter the pytest process."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "keras_shim")


def run(body):
    code = "import sys\nsys.path.insert(0, %r)\nimport numpy as np, torch\n" % SHIM + textwrap.dedent(body)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_batch_dot_axis_rule_is_the_einsum_the_reference_relies_on():
    """K.batch_dot(D (B,NF), A (B,NF,NF), axes=[1,1]) (BS_brain.py:73) = out[b,j] = sum_i D[b,i] A[b,i,j]."""
    run("""
        from keras import backend as K
        g = torch.Generator().manual_seed(0)
        D, A = torch.randn(5, 12, generator=g, dtype=torch.float64), torch.randn(5, 12, 12, generator=g, dtype=torch.float64)
        out = K.batch_dot(D, A, axes=[1, 1])
        assert out.shape == (5, 12)
        assert torch.allclose(out, torch.einsum("bi,bij->bj", D, A), atol=1e-12)
        # and the other shapes the 2.2.4 backend defines: 2-D x 2-D -> (B, 1); 3-D x 3-D default axes
        x, y = torch.randn(4, 7, generator=g), torch.randn(4, 7, generator=g)
        assert torch.allclose(K.batch_dot(x, y, axes=1), (x * y).sum(1, keepdim=True))
        p, q = torch.randn(3, 5, 6, generator=g), torch.randn(3, 6, 2, generator=g)
        assert torch.allclose(K.batch_dot(p, q), torch.bmm(p, q), atol=1e-6)
    """)


def test_huber_loss_is_torchs_mean_huber():
    run("""
        import tensorflow as tf
        g = torch.Generator().manual_seed(1)
        y, q = torch.randn(64, 4, generator=g, dtype=torch.float64) * 3, torch.randn(64, 4, generator=g, dtype=torch.float64) * 3
        ref = torch.nn.functional.huber_loss(q, y, reduction="mean", delta=1.0)
        assert abs(float(tf.losses.huber_loss(y, q)) - float(ref)) < 1e-12
    """)


def test_adam_is_torchs_adam_as_epsilon_vanishes_and_keras_form_otherwise():
    """Keras 2.2.4: p -= lr * sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps).  With eps -> 0 this is torch.optim.Adam;
    with the default eps = 1e-7 it is checked against the closed form for the first step (m = (1-b1) g, v = (1-b2) g^2)."""
    run("""
        import os
        os.environ["KERAS_SHIM_FLOATX"] = "float64"
        from keras.optimizers import Adam
        g = torch.Generator().manual_seed(2)
        p0 = torch.randn(50, generator=g, dtype=torch.float64)
        grads = [torch.randn(50, generator=g, dtype=torch.float64) for _ in range(5)]
        mine = p0.clone()
        opt = Adam(lr=0.001, beta_1=0.5, beta_2=0.999, epsilon=1e-30)
        theirs = p0.clone().requires_grad_(True)
        topt = torch.optim.Adam([theirs], lr=0.001, betas=(0.5, 0.999), eps=1e-30)
        for gr in grads:
            opt.apply([mine], [gr])
            theirs.grad = gr.clone()
            topt.step()
        assert torch.allclose(mine, theirs.detach(), rtol=0, atol=1e-12)
        one = p0.clone()
        k = Adam(lr=0.001, beta_1=0.5, beta_2=0.999)               # epsilon None -> K.epsilon() = 1e-7
        k.apply([one], [grads[0]])
        lr_t = 0.001 * np.sqrt(1 - 0.999) / (1 - 0.5)
        want = p0 - lr_t * (0.5 * grads[0]) / (torch.sqrt(0.001 * grads[0] ** 2) + 1e-7)
        assert torch.allclose(one, want, rtol=0, atol=1e-9)          # hyper-parameters are float32-rounded like K.variable
    """)


def test_dense_glorot_and_model_plumbing():
    run("""
        import keras
        from keras.layers import Input, Dense
        from keras.models import Model
        a = Input(shape=(6,), name="a")
        h = Dense(8, activation="relu")(a)
        o = Dense(3, activation="linear", name="out")(keras.layers.concatenate([h, a]))
        m = Model(inputs=[a], outputs=[o])
        W1, b1, W2, b2 = m.get_weights()
        assert W1.shape == (6, 8) and W2.shape == (14, 3) and not b1.any() and not b2.any()
        assert np.abs(W1).max() <= np.sqrt(6.0 / (6 + 8)) and np.abs(W2).max() <= np.sqrt(6.0 / (14 + 3))   # glorot_uniform limit
        x = np.random.default_rng(0).normal(size=(40, 6))
        want = np.concatenate([np.maximum(x.astype(np.float32) @ W1 + b1, 0), x.astype(np.float32)], 1) @ W2 + b2
        assert np.allclose(m.predict({"a": x}), want, atol=1e-5)                # predict in chunks of 32
        m2 = Model(inputs=[a], outputs=[o])
        m2.set_weights([w + 1 for w in m.get_weights()])
        assert all(np.array_equal(u, v + 1) for u, v in zip(m2.get_weights(), [W1, b1, W2, b2]))   # shared layer objects
        # fit: one batch, loss = mean Huber, History key 'loss' only for a single output
        import tensorflow as tf
        m.compile(optimizer=keras.optimizers.Adam(lr=0.001, beta_1=0.5, beta_2=0.999), loss=lambda t, p: tf.losses.huber_loss(t, p))
        y = np.zeros((40, 3))
        before = torch.as_tensor(m.predict({"a": x}))
        h1 = m.fit({"a": x}, {"out": y}, batch_size=40, epochs=1, verbose=0)
        assert list(h1.history) == ["loss"]
        assert abs(h1.history["loss"][0] - float(torch.nn.functional.huber_loss(before, torch.zeros(40, 3)))) < 1e-5
        assert not np.allclose(m.predict({"a": x}), before.numpy())             # the step moved the weights
        # two batches: the epoch loss is the batch-size-weighted mean of the batch losses
        h2 = m.fit({"a": x}, {"out": y}, batch_size=32, epochs=1, verbose=0)
        assert len(h2.history["loss"]) == 1
        try:
            m.predict({"b": x})
            raise SystemExit("missing key accepted")
        except ValueError:
            pass
    """)


def test_layer_auto_names_follow_keras_uid_rule():
    """The generators map unnamed layers to node slots by their creation order (gnn_layer_1.., dense_1..)."""
    run("""
        from keras.layers import Layer, Dense, Input
        class GNNLayer(Layer):
            pass
        assert [GNNLayer().name, GNNLayer().name, Dense(3).name, Dense(3).name, GNNLayer(name="D1_GNN").name] == \\
            ["gnn_layer_1", "gnn_layer_2", "dense_1", "dense_2", "D1_GNN"]
    """)
