"""CPU: group-table generator (restatement of tools/lvis_analyse.py:11-98)."""
import numpy as np
import torch

from balancedgroupsoftmax_b200.tables import (build_group_tables, load_reference_files, save_reference_files,
                                              synthetic_instance_counts, synthetic_tables)


def test_structure_5bin():
    t = synthetic_tables(1231, seed=0)
    assert t.label2binlabel.shape == (5, 1231) and t.label2binlabel.dtype == np.int64
    assert t.pred_slice.shape == (5, 2)
    assert t.num_logits == 1236                      # 1231 + 5 "others"/bg slots
    assert t.pred_slice[0].tolist() == [0, 2]
    # running (start, len)
    assert np.array_equal(t.pred_slice[1:, 0], np.cumsum(t.pred_slice[:-1, 1]))
    # row 0: bg -> 0, any fg -> 1
    assert t.label2binlabel[0, 0] == 0 and (t.label2binlabel[0, 1:] == 1).all()
    # every fg class sits in exactly one bin with a 1-based contiguous local index
    nz = (t.label2binlabel[1:] > 0).sum(0)
    assert nz[0] == 0 and (nz[1:] == 1).all()
    for g in range(1, 5):
        loc = t.label2binlabel[g][t.label2binlabel[g] > 0]
        assert sorted(loc.tolist()) == list(range(1, len(loc) + 1))
        assert t.pred_slice[g, 1] == len(loc) + 1
        # fg_splits order == local index order
        split = t.fg_splits[g - 1]
        assert np.array_equal(t.label2binlabel[g][split], np.arange(1, len(split) + 1))
    assert sum(len(s) for s in t.fg_splits) == 1230


def test_bin_thresholds():
    counts = {1: 1, 2: 9, 3: 10, 4: 99, 5: 100, 6: 999, 7: 1000, 8: 50000}
    t = build_group_tables(counts, num_classes=9)
    assert [s.tolist() for s in t.fg_splits] == [[1, 2], [3, 4], [5, 6], [7, 8]]
    assert t.pred_slice.tolist() == [[0, 2], [2, 3], [5, 3], [8, 3], [11, 3]]
    assert t.label2binlabel[2].tolist() == [0, 0, 0, 1, 2, 0, 0, 0, 0]


def test_cls2col_inverse_of_merge_scatter():
    t = synthetic_tables(1231, seed=3)
    c2c = t.cls2col()
    assert c2c[0] == 0
    assert len(set(c2c.tolist())) == 1231          # every class fed by a distinct column
    others = set(int(s) for s in t.pred_slice[1:, 0]) | {1}
    assert not (set(c2c.tolist()) & others)         # "others" slots and P(fg) never surface as a class


def test_reference_file_roundtrip(tmp_path):
    t = synthetic_tables(1231, seed=1)
    paths = save_reference_files(t, str(tmp_path))
    l2b = torch.load(paths['label2binlabel'])
    assert l2b.dtype == torch.int64 and tuple(l2b.shape) == (5, 1231)
    t2 = load_reference_files(paths['label2binlabel'], paths['pred_slice'], paths['fg_split'])
    assert np.array_equal(t.label2binlabel, t2.label2binlabel)
    assert np.array_equal(t.pred_slice, t2.pred_slice)
    assert all(np.array_equal(a, b) for a, b in zip(t.fg_splits, t2.fg_splits))


def test_other_bin_counts():
    counts = synthetic_instance_counts(1230, seed=0)
    t2 = build_group_tables(counts, thresholds=(100,))            # 3 bins (2-bin fg variant)
    t8 = build_group_tables(counts, thresholds=(5, 10, 50, 100, 500, 1000, 5000))[0:1] if False else \
        build_group_tables(counts, thresholds=(5, 10, 50, 100, 500, 1000))
    assert t2.num_bins == 3 and t2.num_logits == 1231 + 3
    assert t8.num_bins == 8 and t8.num_logits == 1231 + 8


def test_cli_from_annotation_json(tmp_path):
    """tools/lvis_analyse.py get_cate_gs + get_split from a categories list: file order defines the local indices,
    the written files load back through the head's own loader."""
    import json
    from balancedgroupsoftmax_b200 import tables as T
    cats = [dict(id=3, instance_count=5), dict(id=1, instance_count=50000), dict(id=2, instance_count=12),
            dict(id=5, instance_count=999), dict(id=4, instance_count=7)]
    ann = tmp_path / 'ann.json'
    ann.write_text(json.dumps(dict(categories=cats, annotations=[])))
    assert T.main(['--ann', str(ann), '--out', str(tmp_path / 'out')]) == 0
    t = T.load_reference_files(str(tmp_path / 'out' / 'label2binlabel.pt'), str(tmp_path / 'out' / 'pred_slice_with0.pt'),
                               str(tmp_path / 'out' / 'valsplit.pkl'))
    assert t.pred_slice.tolist() == [[0, 2], [2, 3], [5, 2], [7, 2], [9, 2]]
    assert t.label2binlabel[0].tolist() == [0, 1, 1, 1, 1, 1]
    assert t.label2binlabel[1].tolist() == [0, 0, 0, 1, 2, 0]      # ids 3 then 4, in file order
    assert t.label2binlabel[2].tolist() == [0, 0, 1, 0, 0, 0]
    assert t.label2binlabel[3].tolist() == [0, 0, 0, 0, 0, 1]
    assert t.label2binlabel[4].tolist() == [0, 1, 0, 0, 0, 0]
    assert [s.tolist() for s in t.fg_splits] == [[3, 4], [2], [5], [1]]
    # counted from annotations when the categories carry no instance_count
    ann2 = tmp_path / 'ann2.json'
    ann2.write_text(json.dumps(dict(categories=[dict(id=1), dict(id=2)],
                                    annotations=[dict(category_id=2)] * 11 + [dict(category_id=1)] * 3)))
    assert T.instance_counts_from_annotations(str(ann2)) == {1: 3, 2: 11}
