"""CPU tests of the metric callers (regtr_b200/eval.py) against fixtures produced by the
reference's own est.log writer, 3DMatch benchmark, ModelNet metrics and metric aggregation
(tests/golden/make_golden.py:eval_fixtures on the seeded inputs of tests/golden/eval_inputs.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
import eval_inputs as ei  # noqa: E402
from regtr_b200 import eval as E  # noqa: E402


@pytest.fixture(scope='module')
def fx():
    return load_golden('eval')


@pytest.fixture(scope='module')
def run(tmp_path_factory):
    tmp = tmp_path_factory.mktemp('eval')
    scenes = ei.make_scenes()
    gt_dir = str(tmp / 'gt')
    ei.write_gt(scenes, gt_dir)
    w = E.EstLogWriter(str(tmp / 'log'), '3DMatch')
    for scene, d in scenes.items():
        for src, tgt, T in d['est']:
            batch = {'src_xyz': [None], 'src_path': [f'x/{scene}/cloud_bin_{src}.pth'],
                     'tgt_path': [f'x/{scene}/cloud_bin_{tgt}.pth']}
            w.append_batch(batch, {'pose': torch.from_numpy(T[None, None, :3].copy())})
    return scenes, gt_dir, w.root


def test_est_log_writer_is_byte_identical(run, fx):
    _, _, est_dir = run
    got = open(os.path.join(est_dir, 'scene-a', 'est.log'), 'rb').read()
    assert got == fx['est_log_scene_a'].tobytes()


def test_3dmatch_benchmark_matches_reference(run, fx):
    scenes, gt_dir, est_dir = run
    s, recall, per = E.benchmark_3dmatch(est_dir, gt_dir)
    assert s == fx['bench_str'].tobytes().decode('utf-8')
    assert abs(recall - float(fx['bench_recall'])) < 1e-12
    for scene in scenes:
        assert np.array_equal(per[scene]['flags'], fx[f'flags_{scene}'])
        np.testing.assert_allclose(per[scene]['errors'], fx[f'errors_{scene}'], rtol=1e-9, atol=1e-12, equal_nan=True)
    assert any((per[sc]['flags'] == 2).any() for sc in scenes)       # consecutive pairs present and ignored
    assert any((per[sc]['flags'] == 1).any() for sc in scenes)       # failed registrations present


def test_trajectory_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    traj = rng.normal(size=(5, 4, 4))
    meta = [[i, i + 2, 9 if i != 3 else 0] for i in range(5)]        # entry 3 is dropped (falsy third field)
    E.write_trajectory(traj, meta, str(tmp_path / 't.log'))
    keys, back = E.read_trajectory(str(tmp_path / 't.log'))
    assert keys.shape == (4, 3) and list(keys[:, 0]) == ['0', '1', '2', '4']
    np.testing.assert_allclose(back, traj[[0, 1, 2, 4]], atol=1e-11)


def test_mat2quat_against_scipy():
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(1)
    for _ in range(50):
        M = Rotation.from_rotvec(rng.normal(size=3) * rng.uniform(0, 3.1)).as_matrix()
        q = E.mat2quat(M)
        w = Rotation.from_matrix(M).as_quat()[[3, 0, 1, 2]]
        assert q[0] >= 0 and min(np.abs(q - w).max(), np.abs(q + w).max()) < 1e-12


def test_modelnet_metrics_match_reference(fx):
    data, pred = ei.modelnet_batch()
    met = E.compute_modelnet_metrics(data, pred)
    for k, v in met.items():
        np.testing.assert_allclose(v, fx[f'mn_{k}'], rtol=1e-5, atol=1e-7, err_msg=k)
    for k, v in E.summarize_modelnet_metrics(met).items():
        np.testing.assert_allclose(v, fx[f'mns_{k}'], rtol=1e-5, atol=1e-7, err_msg=k)


def test_metric_aggregation_matches_reference(fx):
    per = [E.compute_metrics({'pose': p}, g) for p, g in ei.pose_batches()]
    agg = E.aggregate_metrics(per, thresh_rot=10, thresh_trans=0.1)
    keys = {k[4:] for k in fx if k.startswith('agg_')}
    assert set(agg) == keys
    for k in keys:
        np.testing.assert_allclose(agg[k].numpy(), fx[f'agg_{k}'], rtol=1e-5, atol=1e-6, err_msg=k)
    assert 0.0 < float(agg['reg_success_final']) <= 1.0


def test_benchmark_loop_end_to_end(tmp_path):
    """Dataset files -> read-ahead stream -> (stub) forward -> est.log -> registration recall: the host loop of
    the reference's test.py with a forward that returns the ground truth pose of every third pair exactly and a
    wrong pose otherwise."""
    import pickle
    from regtr_b200 import data as D
    scenes = ei.make_scenes(seed=5)
    gt_dir = str(tmp_path / 'gt')
    ei.write_gt(scenes, gt_dir)
    rng = np.random.default_rng(0)
    infos = dict(rot=[], trans=[], src=[], tgt=[], overlap=[])
    for scene, d in scenes.items():
        os.makedirs(tmp_path / 'data' / 'test' / scene, exist_ok=True)
        for i in range(d['n_frag']):
            torch.save(rng.normal(size=(40, 3)), tmp_path / 'data' / 'test' / scene / f'cloud_bin_{i}.pth')
        for (i, j), T in zip(d['pairs'], d['gt']):                      # gt.log pair (i, j): src = j, tgt = i
            infos['rot'].append(T[:3, :3]); infos['trans'].append(T[:3, 3:4])
            infos['src'].append(f'test/{scene}/cloud_bin_{j}.pth'); infos['tgt'].append(f'test/{scene}/cloud_bin_{i}.pth')
            infos['overlap'].append(0.5)
    with open(tmp_path / 'info.pkl', 'wb') as f:
        pickle.dump(infos, f)
    ds = D.ThreeDMatchPairs(str(tmp_path / 'data'), str(tmp_path / 'info.pkl'))
    batches = [list(range(i, min(i + 2, len(ds)))) for i in range(0, len(ds), 2)]

    def forward(batch):
        poses = batch['pose'].clone()
        for b, idx in enumerate(batch['idx']):
            if idx % 3:
                poses[b, :, 3] += 1.0                                     # a 1 m error: registration failure
        return {'pose': poses[None].repeat(6, 1, 1, 1)}

    res = E.run_3dmatch_benchmark(D.PairStream(ds, batches, workers=2), forward, str(tmp_path / 'log'), '3DMatch', gt_dir)
    assert 0.25 < res['recall'] < 0.45                                    # every third pair succeeds
    assert abs(float(res['metrics']['reg_success_final']) - np.mean([i % 3 == 0 for i in range(len(ds))])) < 1e-6
    assert 'Mean median RRE' in res['summary']
