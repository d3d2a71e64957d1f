"""CPU tests of the input pipeline (regtr_b200/data.py): file formats of the reference's
3DMatch loader, collate layout, size bucketing, read-ahead stream."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

from regtr_b200 import data as D

REF = '/root/reference'


def _make_dataset(tmp_path, n_pairs=11, seed=0):
    rng = np.random.default_rng(seed)
    infos = dict(rot=[], trans=[], src=[], tgt=[], overlap=[])
    os.makedirs(tmp_path / 'test' / 'scene', exist_ok=True)
    for i in range(n_pairs + 1):
        pts = rng.normal(size=(int(rng.integers(50, 400)), 3))                      # float64, like the real files
        torch.save(pts, tmp_path / 'test' / 'scene' / f'cloud_bin_{i}.pth')
    for i in range(n_pairs):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        infos['rot'].append(q); infos['trans'].append(rng.normal(size=(3, 1)))
        infos['src'].append(f'test/scene/cloud_bin_{i + 1}.pth'); infos['tgt'].append(f'test/scene/cloud_bin_{i}.pth')
        infos['overlap'].append(float(rng.uniform(0.1, 0.9)))
    with open(tmp_path / 'info.pkl', 'wb') as f:
        pickle.dump(infos, f)
    return D.ThreeDMatchPairs(str(tmp_path), str(tmp_path / 'info.pkl')), infos


def test_dataset_items_and_collate(tmp_path):
    ds, infos = _make_dataset(tmp_path)
    assert len(ds) == 11
    it = ds[3]
    assert it['src_xyz'].dtype == torch.float32 and it['src_xyz'].shape[1] == 3
    assert it['pose'].shape == (3, 4) and it['idx'] == 3 and it['src_path'].endswith('cloud_bin_4.pth')
    np.testing.assert_allclose(it['pose'][:, :3].numpy(), infos['rot'][3], atol=1e-6)
    np.testing.assert_allclose(it['pose'][:, 3:].numpy(), infos['trans'][3], atol=1e-6)
    b = D.collate_pair([ds[0], ds[1], ds[2]])
    assert isinstance(b['src_xyz'], list) and len(b['tgt_xyz']) == 3 and b['pose'].shape == (3, 3, 4)
    assert b['overlap_p'].shape == (3,) and b['idx'] == [0, 1, 2]


def test_collate_matches_reference_function(tmp_path):
    if not os.path.isdir(REF):
        pytest.skip('reference not present on this machine')
    import importlib.util                       # the file only imports torch; its package __init__ needs h5py
    spec = importlib.util.spec_from_file_location('ref_collate_functions',
                                                  os.path.join(REF, 'src', 'data_loaders', 'collate_functions.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ref_collate = mod.collate_pair
    ds, _ = _make_dataset(tmp_path)
    items = [ds[i] for i in (4, 0, 7)]
    a, b = D.collate_pair(items), ref_collate(items)
    assert set(a) == set(b)
    for k in a:
        if torch.is_tensor(a[k]):
            assert torch.equal(a[k], b[k])
        else:
            assert len(a[k]) == len(b[k]) and all(
                (torch.equal(x, y) if torch.is_tensor(x) else x == y) for x, y in zip(a[k], b[k]))


def test_reads_the_reference_sample_pairs():
    """The 3DMatch pairs shipped with the reference (data/indoor, demo.py) through the real benchmark pickle."""
    root, info = os.path.join(REF, 'data', 'indoor'), os.path.join(REF, 'src', 'datasets', '3dmatch', 'test_3DMatch_info.pkl')
    if not (os.path.isdir(root) and os.path.isfile(info)):
        pytest.skip('reference data not present on this machine')
    ds = D.ThreeDMatchPairs(root, info)
    assert len(ds) == 1623
    have = [i for i in range(len(ds)) if all(os.path.isfile(os.path.join(root, ds.infos[k][i])) for k in ('src', 'tgt'))]
    assert have, 'no complete pair among the shipped sample clouds'
    it = ds[have[0]]
    raw = torch.load(os.path.join(root, it['src_path']), weights_only=False)
    assert raw.dtype == np.float64 and torch.equal(it['src_xyz'], torch.from_numpy(raw).float())
    assert 5000 < it['src_xyz'].shape[0] < 60000 and it['pose'].shape == (3, 4)


def test_bucket_order_is_a_partition_grouped_by_capacity():
    rng = np.random.default_rng(1)
    sizes = rng.integers(20000, 60000, size=200).tolist()
    batches = D.bucket_order(sizes, batch_size=4, bucket=8192, window=64)
    flat = [i for b in batches for i in b]
    assert sorted(flat) == list(range(200))
    assert all(len(b) <= 4 for b in batches)
    assert max(flat[:64]) < 64 and min(flat[64:128]) >= 64                   # windows keep the dataset order
    cap = lambda i: -(-sizes[i] // 8192)
    assert all(cap(a) <= cap(b) for a, b in zip(flat[:63], flat[1:64]))       # grouped by bucket inside a window
    changes = sum(cap(a) != cap(b) for a, b in zip(flat, flat[1:]))
    naive = sum(cap(a) != cap(b) for a, b in zip(range(199), range(1, 200)))
    assert changes < naive / 3                                                # far fewer graph switches


def test_pair_stream_order_and_errors(tmp_path):
    ds, _ = _make_dataset(tmp_path, n_pairs=23)
    batches = D.bucket_order(ds.sizes(), batch_size=3, bucket=256, window=8)
    got = [b['idx'] for b in D.PairStream(ds, batches, workers=3, depth=2)]
    assert got == [list(b) for b in batches]
    os.remove(tmp_path / 'test' / 'scene' / 'cloud_bin_5.pth')
    with pytest.raises(Exception):
        list(D.PairStream(ds, batches, workers=2, depth=2))


def test_sequence_helpers_match_reference():
    """regtr_b200.seq vs utils/seq_manipulation.py (imported when the reference is present), plus a padded
    round trip."""
    from regtr_b200 import seq as S
    rng = np.random.default_rng(2)
    seqs = [torch.from_numpy(rng.normal(size=(n, 5)).astype(np.float32)) for n in (7, 3, 11, 1)]
    padded, mask, lens = S.pad_sequence(seqs, require_padding_mask=True, require_lens=True)
    assert padded.shape == (11, 4, 5) and mask.shape == (4, 11) and lens == [7, 3, 11, 1]
    assert mask[1, 3:].all() and not mask[1, :3].any() and not mask[2].any()
    back = S.unpad_sequences(padded, lens)
    assert all(torch.equal(a, b) for a, b in zip(back, seqs))
    stacked = torch.cat(seqs)
    src, tgt = S.split_src_tgt(stacked, torch.tensor(lens))
    assert len(src) == 2 and torch.equal(tgt[1], seqs[3])
    if os.path.isdir(REF):
        import importlib.util
        spec = importlib.util.spec_from_file_location('ref_seq', os.path.join(REF, 'src', 'utils', 'seq_manipulation.py'))
        R = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(R)
        for kw in (dict(), dict(require_padding_mask=True, require_lens=True)):
            a, b = S.pad_sequence(seqs, **kw), R.pad_sequence(seqs, **kw)
            assert torch.equal(a[0], b[0]) and (a[1] is None) == (b[1] is None) and a[2] == b[2]
            if a[1] is not None:
                assert torch.equal(a[1], b[1])
        six = torch.from_numpy(rng.normal(size=(6, 11, 4, 5)).astype(np.float32))
        assert all(torch.equal(x, y) for x, y in zip(S.unpad_sequences(six, lens), R.unpad_sequences(six, lens)))
        a, b = S.split_src_tgt(stacked, lens), R.split_src_tgt(stacked, lens)
        assert all(torch.equal(x, y) for x, y in zip(a[0] + a[1], b[0] + b[1]))


def test_compute_overlap_properties(tmp_path):
    rng = np.random.default_rng(4)
    tgt = rng.uniform(0, 1, size=(400, 3))
    src = np.concatenate([tgt[:250] + rng.normal(scale=0.002, size=(250, 3)), rng.uniform(2, 3, size=(80, 3))])
    sm, tm, corr = D.compute_overlap(src, tgt, 0.02)
    assert sm[:250].all() and not sm[250:].any()                 # the shifted copies overlap, the far blob does not
    assert tm[:250].all() and tm[250:].sum() < 40
    s_idx, t_idx = corr
    assert (np.linalg.norm(src[s_idx] - tgt[t_idx], axis=1) < 0.02).all()
    assert (s_idx[t_idx == s_idx] > 0).all() and len(s_idx) > 200     # mutual matches, index 0 excluded (reference quirk)
    # brute-force nearest neighbour agrees
    d = np.linalg.norm(src[:, None] - tgt[None], axis=-1)
    assert np.array_equal(sm, d.min(1) < 0.02)
    ds, infos = _make_dataset(tmp_path)
    ds.overlap_radius = 0.5
    it = ds[2]
    assert it['src_overlap'].dtype == torch.bool and it['src_overlap'].shape[0] == it['src_xyz'].shape[0]
    assert it['correspondences'].shape[0] == 2
    b = D.collate_pair([ds[0], ds[1]])
    assert isinstance(b['correspondences'], list) and len(b['src_overlap']) == 2
