import sys, os, json
sys.path.insert(0, '.')
from freddie_b200 import packed
out = sys.argv[1]
idx = json.load(open(os.path.join(out, 'packed_segment', 'index.json')))
n = 0
for b in idx['batches']:
    ps = packed.PackedSegment(os.path.join(out, 'packed_segment', b['file']))
    for k, (c, t) in enumerate(ps.tints()):
        assert ps.text(k) == open(os.path.join(out, c, 'segment_%s_%d.tsv' % (c, t))).read()
        n += 1
    ps.read_segment()
print('packed segment ok', n)
