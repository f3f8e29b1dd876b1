"""Bucket boundaries for variable-length sequences (reference: lstm/best_buckets.py:1-71):
greedy splitting of the length histogram that maximises the padding area saved per split."""


def calculate_buckets(array, max_length, max_buckets):
    """array: [(user, [items])].  Returns at most max_buckets bucket lengths (unsorted, as the
    reference; callers sort)."""
    hist = {}
    for _, seq in array:
        hist[len(seq)] = hist.get(len(seq), 0) + 1
    running, s = [], 0
    for l in sorted(hist):
        s += hist[l]
        running.append((l, s))                       # (length, #sequences with length <= l)

    def best_point(ll):
        # index i so that splitting into ll[:i+1] | ll[i+1:] saves the most padding
        index, maxv, base = 0, 0, ll[0][1]
        for i, (l, n) in enumerate(ll):
            v = (ll[-1][0] - l) * (n - base)
            if v > maxv:
                maxv, index = v, i
        return index, maxv

    end_index = 0
    for i in range(len(running) - 1, -1, -1):
        if running[i][0] <= max_length:
            end_index = i + 1
            break
    if end_index <= max_buckets:
        return [x[0] for x in running[:end_index]]
    buckets = []
    states = [(running[:end_index], 0, end_index - 1)]      # (segment, gain, split index)
    while len(buckets) < max_buckets:
        k = max(range(len(states)), key=lambda j: (states[j][1], -j))       # first maximum, like arg_max
        seg, _, split = states.pop(k)
        buckets.append(seg[split][0])
        for part in (seg[:split + 1], seg[split + 1:]):
            if part:
                idx, gain = best_point(part)
                states.append((part, gain, idx))
    return buckets
