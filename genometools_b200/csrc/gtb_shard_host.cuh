// gtb_shard_host.cuh -- host side of the sharded job (see gtb_shard.cuh): run_sharded, the
// function every code range executes, and gtb_group, several ranges driven by the threads of one
// process.  Included by gtb_esa.cu after the stage functions.
#pragma once

namespace {

using namespace gtb;

// ---- the one collective: all-gather of a small host block --------------------------------
struct ShardComm {
  int me = 0, world = 1;
  gtb_allgather_fn ag = nullptr;
  void *ctx = nullptr;
  bool ipc = false;               // the other ranges live in other processes (their buffers are mapped, gtb_vmm.cuh)
};

struct SyncHead { int rc; char msg[124]; };

// Every sync point carries the status of the phase before it: if any range failed, all ranges
// leave together with its message (nobody is left waiting at the next sync point).
template <class T>
int sync_gather(gtb_esa *h, ShardComm &c, int rc, const T &mine, std::vector<T> &all)
{
  struct Block { SyncHead head; T body; };
  static_assert(std::is_trivially_copyable<T>::value, "wire format");
  Block b;
  memset(&b, 0, sizeof b);
  b.head.rc = rc;
  if (rc != 0) snprintf(b.head.msg, sizeof b.head.msg, "%.120s", h->err.msg);
  b.body = mine;
  std::vector<Block> blocks((size_t) c.world);
  if (c.ag(c.ctx, &b, sizeof b, blocks.data()) != 0) {
    if (rc == 0) h->err.set("all-gather between the code ranges failed");
    return -1;
  }
  all.resize((size_t) c.world);
  int bad = -1;
  for (int r = 0; r < c.world; r++) {
    all[(size_t) r] = blocks[(size_t) r].body;
    if (blocks[(size_t) r].head.rc != 0 && bad < 0) bad = r;
  }
  if (bad >= 0) {
    if (bad != c.me) h->err.set("code range %d failed: %s", bad, blocks[(size_t) bad].head.msg);
    return -1;
  }
  return 0;
}
struct Nothing { int unused; };
int sync_barrier(gtb_esa *h, ShardComm &c, int rc)
{
  std::vector<Nothing> all;
  return sync_gather(h, c, rc, Nothing{0}, all);
}

// ---- peer pointers ---------------------------------------------------------------------------
// the buffers other ranges read or write: in a job of separate processes they are allocated with
// the virtual-memory API (DevBuf::share_dev) and travel as file descriptors (gtb_vmm.cuh)
constexpr int SHARE_SLOTS = 7;
DevBuf *share_slot(gtb_esa *h, int slot)
{
  DevBuf *t[SHARE_SLOTS] = {&h->kbuf[0], &h->kbuf[1], &h->vbuf[0], &h->vbuf[1], &h->rankwords, &h->trank, &h->leftborder};
  return slot >= 0 && slot < SHARE_SLOTS ? t[slot] : nullptr;
}
int slot_of(gtb_esa *h, const DevBuf *b)
{
  for (int i = 0; i < SHARE_SLOTS; i++) if (share_slot(h, i) == b) return i;
  return -1;
}

// descriptors are received by a thread of their own and kept until the mapping is asked for
void ipc_receiver(gtb_esa *h)
{
  while (!h->ipc_stop.load()) {
    struct pollfd p; p.fd = h->ipc_sock; p.events = POLLIN; p.revents = 0;
    const int r = poll(&p, 1, 100);
    if (r <= 0 || !(p.revents & POLLIN)) continue;
    gtb_esa::PendingFd pf;
    ErrBuf e2;
    if (fd_recv(h->ipc_sock, &pf.msg, &pf.fd, e2) != 0) continue;
    { std::lock_guard<std::mutex> lk(h->ipc_mu); h->pending_fds.push_back(pf); }
    h->ipc_cv.notify_all();
  }
}
void stop_ipc_receiver(gtb_esa *h)
{
  if (h->ipc_thread.joinable()) {
    h->ipc_stop.store(true);
    h->ipc_thread.join();
    h->ipc_stop.store(false);
  }
}

int send_fd_everywhere(gtb_esa *h, const ShardComm &c, int slot, const DevBuf &b)
{
  ErrBuf &err = h->err;
  int fd = -1;
  GTB_TRY(vmm_export_fd(b.mh, &fd, err));
  FdMsg m; m.from = c.me; m.slot = slot; m.alloc_id = b.alloc_id; m.size = b.cap;
  int rc = 0;
  for (int r = 0; r < c.world && rc == 0; r++) {
    if (r == c.me) continue;
    for (int waited = 0;; waited++) {
      rc = fd_send(h->ipc_sock, h->ipc_key.c_str(), r, m, fd, err);
      if (rc != 1) break;
      if (waited > 120000) { err.set("code range %d does not take memory descriptors", r); rc = -1; break; }
      usleep(500);                         // (its receiver thread empties the queue)
    }
  }
  close(fd);
  return rc;
}

int export_ptr(gtb_esa *h, const ShardComm &c, const DevBuf *b, bool present, PeerPtr *out)
{
  ErrBuf &err = h->err;
  memset(out, 0, sizeof *out);
  if (!present || !b || !b->p) return 0;
  out->ptr = (u64) (uintptr_t) b->p;
  out->device = h->device;
  out->valid = 1;
  out->slot = slot_of(h, b);
  out->alloc_id = b->alloc_id; out->size = b->cap;
  if (c.ipc) {
    if (!b->vmm || out->slot < 0) { err.set("internal: exported buffer %d is not shareable", out->slot); return -1; }
    if (h->sent_id[out->slot] != b->alloc_id) {          // a new allocation: the peers need its descriptor
      GTB_TRY(send_fd_everywhere(h, c, out->slot, *b));
      h->sent_id[out->slot] = b->alloc_id;
    }
  }
  return 0;
}

// the address under which this range's GPU reaches a buffer of range `peer`
int resolve_ptr(gtb_esa *h, const ShardComm &c, const PeerPtr &p, int peer, void **out)
{
  ErrBuf &err = h->err;
  *out = nullptr;
  if (!p.valid || p.ptr == 0) return 0;
  if (peer == c.me) { *out = (void *) (uintptr_t) p.ptr; return 0; }
  if (!c.ipc) {
    if (p.device != h->device && !(h->peer_enabled & (1ull << (p.device & 63)))) {
      int can = 0;
      GTB_CUDA(cudaDeviceCanAccessPeer(&can, h->device, p.device));
      if (!can) { err.set("device %d cannot access the memory of device %d (no peer access)", h->device, p.device); return -1; }
      cudaError_t e = cudaDeviceEnablePeerAccess(p.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        err.set("cudaDeviceEnablePeerAccess(%d) failed: %s", p.device, cudaGetErrorString(e)); return -1;
      }
      cudaGetLastError();
      h->peer_enabled |= 1ull << (p.device & 63);
    }
    *out = (void *) (uintptr_t) p.ptr;
    return 0;
  }
  // another process: its allocation is mapped once and kept for the following runs
  const std::pair<int, int> key(peer, p.slot);
  auto it = h->imports.find(key);
  if (it != h->imports.end() && it->second.alloc_id == p.alloc_id) { *out = it->second.ptr; return 0; }
  if (it != h->imports.end()) { vmm_free(it->second.ptr, it->second.mh, it->second.size); h->imports.erase(it); }
  int fd = -1;
  {
    std::unique_lock<std::mutex> lk(h->ipc_mu);
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(120);
    for (;;) {
      for (size_t i = 0; i < h->pending_fds.size(); i++) {
        const FdMsg &m = h->pending_fds[i].msg;
        if (m.from == peer && m.slot == p.slot) {
          if (m.alloc_id == p.alloc_id) fd = h->pending_fds[i].fd; else close(h->pending_fds[i].fd);   // (stale: reallocated since)
          h->pending_fds.erase(h->pending_fds.begin() + (long) i);
          i--;
          if (fd >= 0) break;
        }
      }
      if (fd >= 0) break;
      if (h->ipc_cv.wait_until(lk, deadline) == std::cv_status::timeout) {
        err.set("timed out waiting for the descriptor of buffer %d of code range %d", p.slot, peer);
        return -1;
      }
    }
  }
  gtb_esa::PeerImport imp;
  imp.alloc_id = p.alloc_id; imp.size = (size_t) p.size; imp.ptr = nullptr; imp.mh = 0;
  const int rc = vmm_import(h->device, fd, imp.size, &imp.ptr, &imp.mh, err);
  close(fd);
  GTB_TRY(rc);
  if (getenv("GTB200_SHARD_TRACE")) fprintf(stderr, "[gtb shard %d/%d] mapped buffer %d of range %d (%zu MiB)\n", c.me, c.world, p.slot, peer, imp.size >> 20);
  h->imports[key] = imp;
  *out = imp.ptr;
  return 0;
}

// a job of separate processes: the shareable buffers move to the virtual-memory allocator, the
// descriptor socket of this rank is bound (before the first sync point: everybody can send after it)
int enter_ipc_mode(gtb_esa *h, const ShardComm &c)
{
  ErrBuf &err = h->err;
  for (int i = 0; i < SHARE_SLOTS; i++) share_slot(h, i)->share_dev = h->device;
  const char *k = getenv("GTB200_IPC_KEY");
  if (!k) k = getenv("MASTER_PORT");
  if (!k) k = "job";
  char name[96];
  snprintf(name, sizeof name, "%.40s-w%d", k, c.world);
  if (h->ipc_sock >= 0 && h->ipc_key != name) { stop_ipc_receiver(h); close(h->ipc_sock); h->ipc_sock = -1; }
  if (h->ipc_sock < 0) {
    h->ipc_key = name;
    stop_ipc_receiver(h);
    h->ipc_sock = fd_sock_open(name, c.me, err);
    if (h->ipc_sock < 0) return -1;
    for (int i = 0; i < 8; i++) h->sent_id[i] = 0;
    h->ipc_thread = std::thread(ipc_receiver, h);
  }
  return 0;
}

// gt_suftabparts_new on coarse buckets (host table of ncoarse counts): out4[4p..] = mincode,
// maxcode (fine-grained codes), global offset, width of part p; returns the number of parts made
unsigned coarse_cut(const u64 *cnt, u32 ncoarse, u64 ncodes, unsigned numofparts, u64 *out4)
{
  std::vector<u64> lb((size_t) ncoarse + 2);
  lb[0] = 0;
  for (u32 c = 0; c < ncoarse; c++) lb[c + 1] = lb[c] + cnt[c];
  const u64 total = lb[ncoarse], fine = ncodes / ncoarse;    // fine codes per coarse code
  unsigned np = 0;
  u64 mincode = 0, target = 0;
  const u64 width = total / numofparts, rem = total % numofparts;
  if (numofparts <= 1 || total <= numofparts || ncoarse == 1) {
    out4[0] = 0; out4[1] = ncodes - 1; out4[2] = 0; out4[3] = total; return 1;
  }
  for (unsigned part = 0; part < numofparts && mincode < ncoarse; part++) {
    target += width + (part < rem ? 1 : 0);
    u64 maxcode;
    if (part == numofparts - 1) maxcode = ncoarse - 1;
    else {
      u64 lo = 0, hi = ncoarse;
      while (lo < hi) { const u64 mid = (lo + hi) >> 1; if (lb[mid + 1] < target) lo = mid + 1; else hi = mid; }
      maxcode = lo < mincode ? mincode : lo;
      if (maxcode > ncoarse - 1) maxcode = ncoarse - 1;
    }
    const u64 w = lb[maxcode + 1] - lb[mincode];
    if (w > 0 || part == numofparts - 1) {
      out4[4 * np] = mincode * fine; out4[4 * np + 1] = (maxcode + 1) * fine - 1;
      out4[4 * np + 2] = lb[mincode]; out4[4 * np + 3] = w; np++;
    }
    mincode = maxcode + 1;
  }
  if (np > 0 && out4[4 * (np - 1) + 1] != ncodes - 1) {
    const u64 mc = out4[4 * (np - 1)] / fine;
    out4[4 * (np - 1) + 1] = ncodes - 1;
    out4[4 * (np - 1) + 3] = total - lb[mc];
  }
  unsigned keep = 0;
  for (unsigned p = 0; p < np; p++)
    if (out4[4 * p + 3] > 0) { for (int q = 0; q < 4; q++) out4[4 * keep + q] = out4[4 * p + q]; keep++; }
  if (keep == 0) { out4[0] = 0; out4[1] = ncodes - 1; out4[2] = 0; out4[3] = total; keep = 1; }
  // the kept parts cover ALL codes (a range fills the bucket-table entries of its own codes): the
  // codes of dropped, empty parts go to their neighbours
  out4[0] = 0;
  out4[4 * (keep - 1) + 1] = ncodes - 1;
  for (unsigned p = 1; p < keep; p++) out4[4 * p] = out4[4 * (p - 1) + 1] + 1;
  return keep;
}

struct CountsMsg { u64 counts[MAX_RANGES]; PeerPtr recv; };
struct TiedMsg { u64 M; };
struct SeamMsg { int nonempty, pad; u64 first_key, last_key, nllv; };

// this range took no part of the codes (fewer parts than ranges): an empty, finished run
int finish_empty(gtb_esa *h, unsigned flags)
{
  h->flags = flags;
  // its bucket tables are part of the job's sum: all zero
  h->counted = false; h->lb_own = false;
  GTB_TRY(count_codes(h, h->pl, false));
  h->lb_own = true;
  h->N = 0; h->entries = 0; h->nllv = 0; h->M0 = h->M = 0; h->sa_offset = 0;
  h->first_key = h->last_key = 0;
  h->stats.totallength = h->n; h->stats.specialcharacters = h->S; h->stats.nonspecials = 0;
  h->stats.longest = ~0ull; h->stats.prefixlength = h->pl; h->stats.numofchars = h->K;
  h->in_progress = false;
  return 0;
}

template <bool DNA>
int sharded_body(gtb_esa *h, ShardComm &c, unsigned pl, unsigned flags)
{
  ErrBuf &err = h->err;
  const int me = c.me, world = c.world;
  cudaStream_t st = h->st;
  const u64 n = h->n;
  int rc;

  // ---- 1. count "allreduce": every range counts the coarse codes of its slice of the text
  //         positions, the tables are gathered and summed, every range cuts the same parts ----
  const u64 lo = n * (u64) me / (u64) world, hi = n * (u64) (me + 1) / (u64) world;   // (n < 2^32, world <= 64)
  struct CoarseMsg { u32 cnt[CC_MAXCODES]; };
  static_assert(sizeof(CoarseMsg) == 4 * CC_MAXCODES, "coarse table block");
  CoarseMsg *mine_c = new (std::nothrow) CoarseMsg;
  std::vector<CoarseMsg> all_c;
  if (!mine_c) { err.set("out of host memory"); rc = -1; }
  else {
    memset(mine_c, 0, sizeof *mine_c);
    u32 *dcnt = nullptr; u64 ncnt = 0;
    rc = gtb_esa_coarse_partial(h, pl, lo, hi, &dcnt, &ncnt);
    if (rc == 0 && cudaMemcpyAsync(mine_c->cnt, dcnt, sizeof(u32) * ncnt, cudaMemcpyDeviceToHost, st) != cudaSuccess) { err.set("copy of the coarse counts failed"); rc = -1; }
    if (rc == 0 && cudaStreamSynchronize(st) != cudaSuccess) { err.set("coarse counts failed: %s", cudaGetErrorString(cudaGetLastError())); rc = -1; }
  }
  {
    CoarseMsg dummy; if (!mine_c) memset(&dummy, 0, sizeof dummy);
    const int g = sync_gather(h, c, rc, mine_c ? *mine_c : dummy, all_c);
    delete mine_c;
    GTB_TRY(g);
  }
  h->stats.ms_count = h->ms_count_ext; h->stats.ms_total += h->ms_count_ext; h->ms_count_ext = 0;
  std::vector<u64> total_c((size_t) h->ncoarse, 0);
  for (int r = 0; r < world; r++) for (u32 i = 0; i < h->ncoarse; i++) total_c[i] += all_c[(size_t) r].cnt[i];
  all_c.clear(); all_c.shrink_to_fit();
  u64 out4[4 * MAX_RANGES];
  const int np = (int) coarse_cut(total_c.data(), h->ncoarse, h->ncodes, (unsigned) world, out4);
  const bool active = me < np;
  u64 first_keys[MAX_RANGES];
  for (int r = 0; r < np; r++) first_keys[r] = gtb_code_first_key(h->K, pl, out4[4 * r]);
  h->shard_np = np;
  if (active) GTB_TRY(gtb_esa_set_code_range_known(h, out4[4 * me], out4[4 * me + 1], out4[4 * me + 2], out4[4 * me + 3], me == np - 1));

  // ---- 2. first-level sort of the own range ----
  const char *scan = getenv("GTB200_SHARD_SCAN");          // "filter": every range scans the whole text
  // (with two ranges the redundant scan of the whole text is cheaper than regenerating the keys of
  //  the received positions: measured on 2 B200, c4: 29.9 vs 37.2 ms of key generation + histograms)
  const bool slice_mode = scan ? strcmp(scan, "filter") != 0 : world >= 3;
  if (!slice_mode) {
    rc = 0;
    if (active) rc = timed_stage(h, [&]() -> int { return stage_begin<DNA>(h, flags); });
    else rc = finish_empty(h, flags);
  } else {
    // 2a. the owner's receive buffer (its second value buffer), sizes of the slice's groups
    const u64 width = active ? out4[4 * me + 3] : 0;
    CountsMsg cm; memset(&cm, 0, sizeof cm);
    TextSrc<DNA> src = make_src<DNA>(h, 0, ~0ull);
    h->fmt = choose_fmt(h, pl);
    src.f = h->fmt; src.pos0 = lo;
    h->rw.passes = 0; h->rw.pairs_moved = 0; h->rw.launches = 0;
    rc = timed_stage(h, [&]() -> int {
      if (active) {
        const u64 tailcnt = h->emit_tail ? h->S + 1 : 0;
        for (int i = 0; i < 2; i++) {
          GTB_TRY(h->kbuf[i].ensure(sizeof(u64) * (width + 1), err));
          GTB_TRY(h->vbuf[i].ensure(sizeof(u32) * (width + tailcnt + 1), err));
        }
        GTB_TRY(export_ptr(h, c, &h->vbuf[1], true, &cm.recv));
      }
      PhaseTimer t(h, &h->ext_ms_keygen);
      GTB_TRY(rs_owner_counts(h->rw, st, src, hi - lo, first_keys, np, cm.counts, err));
      t.stop();
      return 0;
    });
    std::vector<CountsMsg> all_m;
    GTB_TRY(sync_gather(h, c, rc, cm, all_m));
    // 2b. the partition pass stores the positions of every group into its owner's buffer
    u64 mytotal = 0, received = 0;
    rc = timed_stage(h, [&]() -> int {
      u64 binbase[MAX_RANGES];
      for (int d = 0; d < np; d++) {
        u64 before = 0;
        for (int r = 0; r < me; r++) before += all_m[(size_t) r].counts[d];
        void *p = nullptr;
        GTB_TRY(resolve_ptr(h, c, all_m[(size_t) d].recv, d, &p));
        if (!p) { err.set("code range %d exported no receive buffer", d); return -1; }
        binbase[d] = (u64) (uintptr_t) (static_cast<u32 *>(p) + before);
        mytotal += cm.counts[d];
      }
      for (int r = 0; r < world; r++) received += active ? all_m[(size_t) r].counts[me] : 0;
      if (active && received != width) {
        err.set("sharded scan: %llu positions for a code range of %llu suffixes", (unsigned long long) received, (unsigned long long) width);
        return -1;
      }
      PhaseTimer t(h, &h->ext_ms_radix);
      GTB_TRY((rs_owner_scatter<TextSrc<DNA>>(h->rw, st, src, hi - lo, np, mytotal, nullptr, nullptr, binbase, err)));
      GTB_CUDA(cudaStreamSynchronize(st));           // the stores have landed in the owners' memory
      t.stop();
      return 0;
    });
    h->ext_pairs = h->rw.pairs_moved; h->ext_launches = h->rw.launches;
    GTB_TRY(sync_barrier(h, c, rc));
    // 2c. the owner regenerates the keys of its positions (text order) and sorts
    rc = 0;
    if (active) {
      rc = timed_stage(h, [&]() -> int {
        // (as in the single-range sort: when tails are rare the keys that carry one are sorted apart and
        //  the first-level sort runs over the symbol digits only)
        const bool tail_last = received > 0 && fmt_tail_digit_alone(h->fmt) && tails_few(h, h->fmt.m);
        const u64 nw = (received + 31) >> 5;
        if (tail_last) GTB_TRY(h->nearbits.ensure(sizeof(u32) * (nw + 2), err));
        if (received > 0) {
          PhaseTimer t(h, &h->ext_ms_keygen);
          k_keys_from_positions<DNA><<<grid_for(received, 256), 256, 0, st>>>(make_src<DNA>(h, 0, ~0ull), h->vbuf[1].as<u32>(), received,
              h->kbuf[1].as<u64>(), tail_last ? h->nearbits.as<u32>() : nullptr);
          GTB_LAUNCH_CHECK();
          h->stats.kernel_launches++;
          t.stop();
        }
        PairSrc ext{h->kbuf[1].as<u64>(), h->vbuf[1].as<u32>()};
        if (!tail_last) return stage_begin<DNA>(h, flags, &ext, received);
        u64 nt = 0;
        u32 *tileoff = nullptr;
        GTB_TRY(device_scan_u32(h, h->nearbits.as<u32>(), nullptr, nw, 1, &tileoff, &nt));
        u64 klo, khi;
        code_range_to_keys(h, &klo, &khi);
        TailSrc tsrc{nullptr, nullptr, klo, khi};
        if (nt > 0) {
          GTB_TRY(h->misc.ensure(256, err));
          GTB_TRY(h->sendidx.ensure(sizeof(u32) * nt, err));
          for (int i = 0; i < 2; i++) {
            GTB_TRY(h->tailkeys[i].ensure(sizeof(u64) * nt, err));
            GTB_TRY(h->tailpos[i].ensure(sizeof(u32) * nt, err));
          }
          // list indices of the tail keys ascending (= text order), their pairs, then stably by the tail
          k_emit_special_tail<<<(unsigned) div_up(nw, SC_TILE), SC_NT, 0, st>>>(h->nearbits.as<u32>(), nw, received, tileoff,
              h->sendidx.as<u32>(), nullptr, 0, reinterpret_cast<unsigned long long *>(h->misc.as<u64>() + 24));
          GTB_LAUNCH_CHECK();
          k_gather_pairs<<<grid_for(nt, 256), 256, 0, st>>>(h->sendidx.as<u32>(), nt, ext.keys, ext.vals,
                                                            h->tailkeys[1].as<u64>(), h->tailpos[1].as<u32>());
          GTB_LAUNCH_CHECK();
          h->stats.kernel_launches += 2;
          PassPlan tp; tp.npass = 0; tp.padded = false;
          plan_add_bits(tp, h->fmt.sh, h->fmt.sh + h->fmt.tb);
          PairSrc ps{h->tailkeys[1].as<u64>(), h->tailpos[1].as<u32>()};
          u64 *tk[2] = {h->tailkeys[0].as<u64>(), h->tailkeys[1].as<u64>()};
          u32 *tv[2] = {h->tailpos[0].as<u32>(), h->tailpos[1].as<u32>()};
          int tres = 0; u64 tout = 0;
          h->rw.passes = 0; h->rw.pairs_moved = 0; h->rw.launches = 0;
          GTB_TRY(radix_sort(h->rw, st, ps, nt, tk, tv, tp, &tres, &tout, err));
          if (tout != nt) { err.set("internal: the tail keys lost elements"); return -1; }
          tsrc.keys = tk[tres]; tsrc.vals = tv[tres];
        }
        return stage_begin<DNA>(h, flags, &ext, received, &tsrc, nt);
      });
    } else rc = finish_empty(h, flags);
  }

  // ---- 3. refinement of the ties in lock step; ranks of foreign positions are read from the
  //         owner's rank map in peer memory ----
  std::vector<TiedMsg> tied;
  GTB_TRY(sync_gather(h, c, rc, TiedMsg{active ? h->M : 0}, tied));
  auto any_tied = [&]() { for (auto &t : tied) if (t.M > 0) return true; return false; };
  if (any_tied()) {
    RankView view; memset(&view, 0, sizeof view);
    rc = 0;
    if (active) rc = timed_stage(h, [&]() -> int {
      GTB_TRY(build_ranks<DNA>(h));
      view.has_map = 1; view.N = h->N; view.sa_offset = h->sa_offset;
      view.own_last = (h->lb_own && !h->counted) ? h->maxcode : ~0ull;
      GTB_TRY(export_ptr(h, c, &h->kbuf[h->res], true, &view.keys));
      GTB_TRY(export_ptr(h, c, &h->vbuf[h->res], true, &view.sa));
      GTB_TRY(export_ptr(h, c, &h->rankwords, true, &view.rw));
      GTB_TRY(export_ptr(h, c, &h->trank, h->M0 > 0, &view.trank));
      GTB_TRY(export_ptr(h, c, &h->leftborder, true, &view.lb));
      GTB_CUDA(cudaStreamSynchronize(st));           // the map is complete before anybody reads it
      return 0;
    });
    std::vector<RankView> views;
    GTB_TRY(sync_gather(h, c, rc, view, views));
    rc = 0;
    if (active) rc = timed_stage(h, [&]() -> int {
      PeerTableDev *pt = new (std::nothrow) PeerTableDev;
      if (!pt) { err.set("out of host memory"); return -1; }
      memset(pt, 0, sizeof *pt);
      pt->n = np; pt->mine = me;
      int r2 = 0;
      for (int r = 0; r < np && r2 == 0; r++) {
        pt->first_key[r] = r == 0 ? 0ull : first_keys[r];
        const RankView &v = views[(size_t) r];
        void *p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        const PeerPtr *pp[5] = {&v.keys, &v.sa, &v.rw, &v.trank, &v.lb};
        for (int i = 0; i < 5 && r2 == 0; i++) r2 = resolve_ptr(h, c, *pp[i], r, &p[i]);
        pt->m[r].keys = static_cast<const u64 *>(p[0]); pt->m[r].sa = static_cast<const u32 *>(p[1]);
        pt->m[r].rw = static_cast<const uint4 *>(p[2]); pt->m[r].trank = static_cast<const u32 *>(p[3]);
        pt->m[r].lb = static_cast<const u32 *>(p[4]);
        pt->m[r].N = v.N; pt->m[r].sa_offset = v.sa_offset; pt->m[r].own_last = v.own_last;
      }
      if (r2 == 0 && h->peertab.ensure(sizeof *pt, err) != 0) r2 = -1;
      if (r2 == 0 && cudaMemcpyAsync(h->peertab.p, pt, sizeof *pt, cudaMemcpyHostToDevice, st) != cudaSuccess) { err.set("upload of the peer table failed"); r2 = -1; }
      if (r2 == 0 && cudaStreamSynchronize(st) != cudaSuccess) { err.set("upload of the peer table failed"); r2 = -1; }
      delete pt;
      return r2;
    });
    int rounds = 0;
    const bool trace = getenv("GTB200_SHARD_TRACE") != nullptr;
    for (;;) {
      const float ms_before = h->stats.ms_doubling;
      const u64 m_before = active ? h->M : 0;
      // (a) sort keys from the ranks as they stand after the previous round -- peer reads
      if (rc == 0 && active && h->M > 0) rc = timed_stage(h, [&]() -> int {
        if (h->round >= 62) { err.set("internal: prefix doubling did not converge"); return -1; }
        PhaseTimer t(h, &h->stats.ms_doubling);
        GTB_TRY(h->sendidx.ensure(sizeof(u32) * h->M, err));   // queue of the partners that need a search
        unsigned int *qcount = reinterpret_cast<unsigned int *>(h->misc.as<u64>() + 20);
        GTB_CUDA(cudaMemsetAsync(qcount, 0, sizeof(unsigned int), st));
        k_build_dkeys_peer<DNA><<<grid_for(h->M, 256), 256, 0, st>>>(make_rankmap<DNA>(h), h->peertab.as<PeerTableDev>(),
            h->upos[h->cur].as<u32>(), h->ugrp[h->cur].as<u32>(), h->M, h->depth[h->round], h->dkeys.as<u64>(),
            h->sendidx.as<u32>(), qcount, (h->isa_round < (unsigned) h->opt_pairs_by_text) ? 1 : 0);
        GTB_LAUNCH_CHECK();
        h->isa_round++;
        k_build_dkeys_peer_search<DNA><<<grid_for(h->M, 256, 148u * 8u), 256, 0, st>>>(make_rankmap<DNA>(h),
            h->peertab.as<PeerTableDev>(), h->upos[h->cur].as<u32>(), h->ugrp[h->cur].as<u32>(), h->depth[h->round],
            h->dkeys.as<u64>(), h->sendidx.as<u32>(), qcount);
        GTB_LAUNCH_CHECK();
        h->stats.kernel_launches += 2;
        GTB_CUDA(cudaStreamSynchronize(st));
        t.stop();
        return 0;
      });
      const float ms_keys = h->stats.ms_doubling - ms_before;
      GTB_TRY(sync_barrier(h, c, rc));                // every range has read: the maps may change
      // (b) sort, write the refined order and the new ranks of the own suffixes
      rc = 0;
      if (active) rc = timed_stage(h, [&]() -> int {
        h->depth[h->round + 1] = 2 * h->depth[h->round];
        if (h->M == 0) { h->round++; return 0; }
        PhaseTimer t(h, &h->stats.ms_doubling);
        GTB_TRY(round_sort_apply<DNA>(h, 0ull, h->bits_lo, false));
        t.stop();
        return 0;
      });
      if (trace) fprintf(stderr, "[gtb shard %d/%d] round %d depth %llu: tied %llu -> %llu, keys from peer ranks %.3f ms, sort+apply %.3f ms\n",
                         me, world, rounds, (unsigned long long) h->depth[h->round > 0 ? h->round - 1 : 0], (unsigned long long) m_before,
                         (unsigned long long) (active ? h->M : 0), ms_keys, h->stats.ms_doubling - ms_before - ms_keys);
      GTB_TRY(sync_gather(h, c, rc, TiedMsg{active ? h->M : 0}, tied));   // every range has written
      if (!any_tied()) break;
      if (++rounds > 64) { err.set("prefix doubling across ranges did not converge"); return -1; }
    }
  }

  // ---- 4. lcp of the deep pairs, stats; seams between neighbouring ranges ----
  rc = 0;
  if (active) rc = timed_stage(h, [&]() -> int { return stage_end<DNA>(h); });
  if (rc == 0) h->ran = true;
  SeamMsg sm; memset(&sm, 0, sizeof sm);
  if (rc == 0 && active) { sm.nonempty = h->N > 0; sm.first_key = h->first_key; sm.last_key = h->last_key; sm.nllv = h->nllv; }
  std::vector<SeamMsg> seams;
  GTB_TRY(sync_gather(h, c, rc, sm, seams));
  h->llv_before = 0;
  bool have_prev = false; u64 prev = 0;
  for (int r = 0; r < me; r++) {
    h->llv_before += seams[(size_t) r].nllv;
    if (seams[(size_t) r].nonempty) { have_prev = true; prev = seams[(size_t) r].last_key; }
  }
  if (active && h->N > 0 && have_prev && (flags & GTB_WANT_LCP)) GTB_TRY(gtb_esa_fix_seam(h, prev));
  return 0;
}

int run_sharded(gtb_esa *h, ShardComm &c, unsigned pl, unsigned flags)
{
  ErrBuf &err = h->err;
  if (c.world < 1 || c.world > MAX_RANGES || c.me < 0 || c.me >= c.world) { err.set("bad rank %d of %d code ranges (1..%d)", c.me, c.world, MAX_RANGES); return -1; }
  if (c.world == 1) {                     // one range: the plain run
    h->full_range = true; h->range_given = false; h->emit_tail = 1; h->llv_before = 0; h->shard_np = 1;
    return gtb_esa_run(h, pl, flags);
  }
  int rc = 0;
  if (pl == 0) { err.set("a sharded run needs prefixlength >= 1"); rc = -1; }
  if (rc == 0) rc = check_run_args(h, pl, flags | GTB_WANT_BCK);
  if (rc == 0 && cudaSetDevice(h->device) != cudaSuccess) { err.set("cudaSetDevice(%d) failed", h->device); rc = -1; }
  if (rc == 0 && c.ipc) rc = enter_ipc_mode(h, c);
  h->ext_ms_keygen = 0; h->ext_ms_radix = 0; h->ext_pairs = 0; h->ext_launches = 0;
  GTB_TRY(sync_barrier(h, c, rc));
  return h->dna ? sharded_body<true>(h, c, pl, flags | GTB_WANT_BCK) : sharded_body<false>(h, c, pl, flags | GTB_WANT_BCK);
}

// ---- all ranges in one process: the all-gather is a shared buffer and a barrier ---------------
struct LocalComm {
  int world = 1;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  u64 gen = 0;
  std::vector<unsigned char> buf;
  void barrier()
  {
    std::unique_lock<std::mutex> lk(mu);
    const u64 g = gen;
    if (++arrived == world) { arrived = 0; gen++; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
  int allgather(int me, const void *mine, size_t bytes, void *all)
  {
    {
      std::lock_guard<std::mutex> lk(mu);
      if (buf.size() < bytes * (size_t) world) buf.resize(bytes * (size_t) world);   // (nobody reads: see the last barrier)
      memcpy(buf.data() + bytes * (size_t) me, mine, bytes);
    }
    barrier();
    memcpy(all, buf.data(), bytes * (size_t) world);
    barrier();
    return 0;
  }
};
struct LocalCtx { LocalComm *comm; int me; };
int local_allgather(void *ctx, const void *mine, size_t bytes, void *all)
{
  LocalCtx *lc = static_cast<LocalCtx *>(ctx);
  return lc->comm->allgather(lc->me, mine, bytes, all);
}

} // namespace

struct gtb_group {
  ErrBuf err;
  std::vector<gtb_esa *> hs;
  LocalComm comm;
  bool ran = false, bck_merged = false;
  unsigned pl = 0, flags = 0;
  gtb_stats stats;
};

namespace {

// run `body(i)` for every range on a thread of its own; the first error message wins
template <class F>
int group_parallel(gtb_group *g, F body)
{
  const int n = (int) g->hs.size();
  std::vector<int> rcs((size_t) n, 0);
  std::vector<std::thread> th;
  th.reserve((size_t) n);
  // every range needs its thread (the ranges wait for each other): if the threads cannot all be had,
  // none is started
  std::mutex gate_mu; std::condition_variable gate_cv; int gate = 0;     // 0 wait, 1 go, -1 give up
  try {
    for (int i = 1; i < n; i++) th.emplace_back([&, i] {
      { std::unique_lock<std::mutex> lk(gate_mu); gate_cv.wait(lk, [&] { return gate != 0; }); if (gate < 0) return; }
      rcs[(size_t) i] = body(i);
    });
  } catch (const std::exception &e) {
    { std::lock_guard<std::mutex> lk(gate_mu); gate = -1; }
    gate_cv.notify_all();
    for (auto &t : th) t.join();
    g->err.set("could not start one thread per code range: %s", e.what());
    return -1;
  }
  { std::lock_guard<std::mutex> lk(gate_mu); gate = 1; }
  gate_cv.notify_all();
  rcs[0] = body(0);
  for (auto &t : th) t.join();
  for (int i = 0; i < n; i++)
    if (rcs[(size_t) i] != 0) {
      snprintf(g->err.msg, sizeof g->err.msg, "%s", g->hs[(size_t) i]->err.msg[0] ? g->hs[(size_t) i]->err.msg : "a code range failed");
      return -1;
    }
  return 0;
}

// the first handle on each device holds the input, the others borrow it
template <class F>
int group_set_input(gtb_group *g, F upload)
{
  const int n = (int) g->hs.size();
  std::vector<int> owner((size_t) n);
  for (int i = 0; i < n; i++) {
    owner[(size_t) i] = i;
    for (int j = 0; j < i; j++) if (g->hs[(size_t) j]->device == g->hs[(size_t) i]->device) { owner[(size_t) i] = j; break; }
  }
  for (int i = 0; i < n; i++) if (owner[(size_t) i] != i) return_borrowed_input(g->hs[(size_t) i]);   // (the owners take a new input)
  GTB_TRY(group_parallel(g, [&](int i) -> int { return owner[(size_t) i] == i ? upload(g->hs[(size_t) i]) : 0; }));
  for (int i = 0; i < n; i++)
    if (owner[(size_t) i] != i && gtb_esa_share_input(g->hs[(size_t) i], g->hs[(size_t) owner[(size_t) i]]) != 0) {
      snprintf(g->err.msg, sizeof g->err.msg, "%s", g->hs[(size_t) i]->err.msg);
      return -1;
    }
  g->ran = false; g->bck_merged = false;
  return 0;
}

// the bucket table of the job = sum of the ranges' tables, gathered on the first range through
// peer memory
int group_merge_bck(gtb_group *g)
{
  if (g->bck_merged || g->hs.size() == 1) { g->bck_merged = true; return 0; }
  gtb_esa *root = g->hs[0];
  ErrBuf &err = g->err;
  GTB_CUDA(cudaSetDevice(root->device));
  ShardComm c; c.me = 0; c.world = (int) g->hs.size(); c.ipc = false;
  for (size_t i = 1; i < g->hs.size(); i++) {
    gtb_esa *o = g->hs[i];
    if (!o->lb_own && !o->counted) continue;            // a range without codes
    GTB_CUDA(cudaSetDevice(o->device));
    GTB_CUDA(cudaStreamSynchronize(o->st));
    GTB_CUDA(cudaSetDevice(root->device));
    const DevBuf *src[3] = {&o->leftborder, &o->csc, &o->dist};
    DevBuf *dst[3] = {&root->leftborder, &root->csc, &root->dist};
    const u64 cnt[3] = {root->ncodes + 1, root->nspecialcodes, root->ndist};
    for (int t = 0; t < 3; t++) {
      if (cnt[t] == 0) continue;
      PeerPtr pp; memset(&pp, 0, sizeof pp);
      pp.ptr = (u64) (uintptr_t) src[t]->p; pp.device = o->device; pp.valid = src[t]->p != nullptr;
      void *p = nullptr;
      if (resolve_ptr(root, c, pp, (int) i, &p) != 0) { snprintf(err.msg, sizeof err.msg, "%s", root->err.msg); return -1; }
      k_add_u32<<<grid_for(cnt[t], 256), 256, 0, root->st>>>(dst[t]->as<u32>(), static_cast<const u32 *>(p), cnt[t]);
      GTB_LAUNCH_CHECK();
    }
  }
  GTB_CUDA(cudaStreamSynchronize(root->st));
  root->counted = true;                                 // (the first range now holds the whole table)
  g->bck_merged = true;
  return 0;
}

} // namespace
