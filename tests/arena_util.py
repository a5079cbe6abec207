"""Host-side walks over the pooled MCTS arena (test infrastructure): the tree of one game as arrays,
with the structural invariants every kernel must keep."""
import numpy as np
import torch


def game_pages(arena, g):
    sel = int(arena.list_sel[g])
    n = int(arena.n_pages_game[g])
    return arena.page_lists[sel, g, :n].cpu().numpy().astype(np.int64)


def walk_tree(arena, g, check=True):
    """Breadth-first walk from game g's root.  Returns dict(index, visits, value, prior, parent, move, n_child,
    first_child, depth) over the reachable nodes and asserts (check=True): child runs lie inside ONE page owned by the
    game, children point back to their parent, N(node) = 1 + sum N(children) for expanded nodes, the node count equals
    arena.n_nodes[g], no slot is reached twice."""
    shift = arena.page_shift
    pages = game_pages(arena, g)
    owned = set(int(p) for p in pages)
    # pull the game's pages to the host once
    idx = (torch.from_numpy(pages).to(arena.device).view(-1, 1) << shift) + torch.arange(1 << shift, device=arena.device).view(1, -1)
    nodes = arena.nodes[idx.view(-1)].cpu().numpy()
    links = arena.links[idx.view(-1)].cpu().numpy()
    slot_of_page = {int(p): i for i, p in enumerate(pages)}

    def local(i):
        return (slot_of_page[int(i) >> shift] << shift) + (int(i) & ((1 << shift) - 1))

    root = int(arena.root[g])
    assert (root >> shift) in owned
    order, depth = [root], [0]
    seen = {root}
    head = 0
    while head < len(order):
        i = order[head]
        li = local(i)
        word = int(links[li, 1])
        nc, fc = word >> 16, int(nodes[li, 3])
        if nc > 0:
            if check:
                assert (fc >> shift) == ((fc + nc - 1) >> shift), "child run straddles a page"
                assert (fc >> shift) in owned, "child run in a page the game does not own"
            kids = range(fc, fc + nc)
            if check:
                ks = [local(k) for k in kids]
                assert all(int(links[k, 0]) == i for k in ks), "child does not point back to its parent"
                assert int(nodes[li, 0]) == 1 + sum(int(nodes[k, 0]) for k in ks) or head == 0, "visit sum"
            for k in kids:
                assert k not in seen
                seen.add(k)
                order.append(k)
                depth.append(depth[head] + 1)
        head += 1
    li = np.array([local(i) for i in order], dtype=np.int64)
    word = links[li, 1].astype(np.int64)
    out = dict(index=np.array(order), visits=nodes[li, 0], value=nodes[li, 1].view(np.float32), prior=nodes[li, 2].view(np.float32),
               first_child=nodes[li, 3], parent=links[li, 0], move=((word & 0xFFFF) ^ 0x8000) - 0x8000, n_child=word >> 16,
               depth=np.array(depth))
    if check:
        assert len(order) == int(arena.n_nodes[g]), (len(order), int(arena.n_nodes[g]))
        assert int(links[local(root), 0]) == -1
    return out


def pool_accounting(arena):
    """Every page is either in the free ring or owned by exactly one game."""
    st = arena.pool_stats()
    ctl = arena.pool_ctl.cpu().numpy()
    head, tail = int(ctl[0]), int(ctl[1])
    ring = arena.free_ring.cpu().numpy()
    free = [int(ring[i % arena.n_pages]) for i in range(head, tail)]
    owned = []
    for g in range(arena.n_games):
        owned.extend(int(p) for p in game_pages(arena, g))
    allp = free + owned
    assert len(allp) == arena.n_pages, (len(free), len(owned), arena.n_pages)
    assert len(set(allp)) == arena.n_pages, "a page is owned twice or both free and owned"
    return st
