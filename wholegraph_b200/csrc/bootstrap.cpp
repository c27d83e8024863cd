/*
 * Control-plane bootstrap for one NVSwitch box: a star of abstract AF_UNIX stream sockets for set-up and file
 * descriptors, plus a shared-memory mailbox (memfd handed out over the star) that carries every small collective
 * afterwards -- barrier, count matrices, argument checks -- in a few microseconds instead of 2(N-1) socket round trips.
 *
 * Replaces, for the single-box scope of this build, the reference's NCCL-based host collectives
 * (cpp/src/wholememory/nccl_comms.cpp:210-232 host_allgather, :383-407 host_alltoall, :82-86
 * barrier) and its AF_UNIX SCM_RIGHTS file-descriptor exchange
 * (cpp/src/wholememory/memory_handle.cpp:746-943).  One mechanism carries both bytes and fds,
 * needs no GPU, no NCCL and no temp directory, so the control plane is testable on a CPU box.
 */
#include "wm_internal.hpp"

#include <algorithm>
#include <atomic>
#include <errno.h>
#include <fcntl.h>
#include <immintrin.h>
#include <poll.h>
#include <sched.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

namespace wm {

namespace {

constexpr int kConnectTimeoutMs = 120000;

/* A collective whose peer never answers (a rank that died without closing its socket cannot happen on one box, but a
 * rank that took a different code path can) must end in an error, not in a hang: every blocking receive gives up after
 * WG_BOOTSTRAP_TIMEOUT_S seconds when that variable is set.  Default: wait forever, like the reference's NCCL control plane --
 * a rank may legitimately sit in a barrier for a long time while another one loads a dataset. */
int recv_timeout_ms()
{
  static const int ms = [] {
    const char* v = getenv("WG_BOOTSTRAP_TIMEOUT_S");
    long s        = (v && *v) ? atol(v) : 0;
    if (s <= 0) return -1;
    return (int)std::min<long>(s, 2000000) * 1000;
  }();
  return ms;
}

/* waits until fd is readable; throws on timeout */
void wait_readable(int fd)
{
  const int limit = recv_timeout_ms();
  for (;;) {
    pollfd pfd{fd, POLLIN, 0};
    int r = ::poll(&pfd, 1, limit);
    if (r > 0) return;
    if (r < 0 && errno == EINTR) continue;
    if (r == 0)
      WM_THROW(WHOLEMEMORY_COMMUNICATION_ERROR, "bootstrap: no answer from a peer rank within %d s (ranks issued different collectives?)",
               limit / 1000);
    WM_THROW(WHOLEMEMORY_COMMUNICATION_ERROR, "bootstrap poll failed: %s", strerror(errno));
  }
}

socklen_t make_addr(const wholememory_unique_id_t& uid, sockaddr_un* addr)
{
  memset(addr, 0, sizeof(*addr));
  addr->sun_family = AF_UNIX;
  /* abstract namespace: sun_path[0] == 0; name = "wgb200-" + 24 hex chars of the id */
  char* p = addr->sun_path + 1;
  int n   = snprintf(p, sizeof(addr->sun_path) - 1, "wgb200-");
  for (int i = 0; i < 12; ++i) n += snprintf(p + n, sizeof(addr->sun_path) - 1 - n, "%02x", (unsigned char)uid.internal[i]);
  return (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
}

int64_t now_ms()
{
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (int64_t)ts.tv_sec * 1000 + ts.tv_nsec / 1000000;
}

/* Who is on the other end?  Abstract sockets carry no file permissions, so the kernel's view of the peer decides:
 * only processes of our own user may join (they could ptrace us anyway); anything else is dropped. */
bool peer_is_our_user(int sock)
{
  ucred cred{};
  socklen_t len = sizeof(cred);
  if (::getsockopt(sock, SOL_SOCKET, SO_PEERCRED, &cred, &len) != 0 || len != sizeof(cred)) return false;
  return cred.uid == ::geteuid();
}

/* receive exactly n bytes before `deadline` (CLOCK_MONOTONIC ms); false on timeout, error or EOF */
bool recv_before(int fd, void* p, size_t n, int64_t deadline)
{
  char* c = static_cast<char*>(p);
  while (n > 0) {
    int remaining = (int)std::max<int64_t>(0, deadline - now_ms());
    pollfd pfd{fd, POLLIN, 0};
    int r = ::poll(&pfd, 1, remaining);
    if (r < 0 && errno == EINTR) continue;
    if (r <= 0) return false;
    ssize_t got = ::recv(fd, c, n, 0);
    if (got < 0 && errno == EINTR) continue;
    if (got <= 0) return false;
    c += got;
    n -= (size_t)got;
  }
  return true;
}

/* What a joining rank says first: its rank and a 16-byte secret taken from unique-id bytes that never appear in the
 * socket name (which is visible in /proc/net/unix).  A local process that merely guesses the name cannot join. */
struct hello {
  int32_t rank;
  unsigned char secret[16];
};
void fill_secret(const wholememory_unique_id_t& uid, unsigned char* out) { memcpy(out, uid.internal + 16, 16); }

}  // namespace

/* ---- shared-memory mailbox ----
 * One slot per rank and per parity bank: { sequence number, payload }.  Collective number k (1, 2, ...) uses bank k & 1:
 * a rank writes its payload, publishes seq = k with release order, then waits (acquire) for seq == k in every other
 * rank's slot and copies the payloads out.  Two banks are enough without a second rendezvous: a rank can only START
 * collective k + 2 (the next user of this bank) after it FINISHED k + 1, i.e. after it saw every rank's seq == k + 1,
 * which every rank publishes only after it finished reading collective k. */
struct bootstrap::mailbox {
  static constexpr size_t kPayload = 1024 - sizeof(std::atomic<uint64_t>);
  struct alignas(128) slot {
    std::atomic<uint64_t> seq;
    unsigned char data[kPayload];
  };
  static_assert(sizeof(slot) == 1024, "slot layout");
  slot* slots     = nullptr; /* [2][size] */
  size_t bytes    = 0;
  uint64_t round  = 0;
};

void bootstrap::send_all(int fd, const void* p, size_t n)
{
  const char* c = static_cast<const char*>(p);
  while (n > 0) {
    ssize_t w = ::send(fd, c, n, MSG_NOSIGNAL);
    if (w < 0) {
      if (errno == EINTR) continue;
      WM_THROW(WHOLEMEMORY_COMMUNICATION_ERROR, "bootstrap send failed: %s", strerror(errno));
    }
    c += w;
    n -= (size_t)w;
  }
}

void bootstrap::recv_all(int fd, void* p, size_t n)
{
  char* c = static_cast<char*>(p);
  while (n > 0) {
    wait_readable(fd);
    ssize_t r = ::recv(fd, c, n, 0);
    if (r < 0) {
      if (errno == EINTR) continue;
      WM_THROW(WHOLEMEMORY_COMMUNICATION_ERROR, "bootstrap recv failed: %s", strerror(errno));
    }
    if (r == 0) WM_THROW(WHOLEMEMORY_COMMUNICATION_ERROR, "bootstrap peer closed the connection");
    c += r;
    n -= (size_t)r;
  }
}

/* one byte of payload says whether an fd rides along */
void bootstrap::send_fd(int sock, int fd)
{
  char tag = fd >= 0 ? 1 : 0;
  if (fd < 0) {
    send_all(sock, &tag, 1);
    return;
  }
  msghdr msg{};
  iovec iov{&tag, 1};
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  msg.msg_iov        = &iov;
  msg.msg_iovlen     = 1;
  msg.msg_control    = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  cmsghdr* c         = CMSG_FIRSTHDR(&msg);
  c->cmsg_level      = SOL_SOCKET;
  c->cmsg_type       = SCM_RIGHTS;
  c->cmsg_len        = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(c), &fd, sizeof(int));
  for (;;) {
    ssize_t w = ::sendmsg(sock, &msg, MSG_NOSIGNAL);
    if (w < 0 && errno == EINTR) continue;
    if (w != 1) WM_THROW(WHOLEMEMORY_COMMUNICATION_ERROR, "bootstrap sendmsg(fd) failed: %s", strerror(errno));
    break;
  }
}

int bootstrap::recv_fd(int sock)
{
  char tag = 0;
  msghdr msg{};
  iovec iov{&tag, 1};
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  msg.msg_iov        = &iov;
  msg.msg_iovlen     = 1;
  msg.msg_control    = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  for (;;) {
    wait_readable(sock);
    ssize_t r = ::recvmsg(sock, &msg, MSG_CMSG_CLOEXEC);
    if (r < 0 && errno == EINTR) continue;
    if (r != 1) WM_THROW(WHOLEMEMORY_COMMUNICATION_ERROR, "bootstrap recvmsg(fd) failed: %s", strerror(errno));
    break;
  }
  if (tag == 0) return -1;
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  if (c == nullptr || c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS)
    WM_THROW(WHOLEMEMORY_COMMUNICATION_ERROR, "bootstrap expected a file descriptor, got none");
  int fd = -1;
  memcpy(&fd, CMSG_DATA(c), sizeof(int));
  return fd;
}

bootstrap::bootstrap(const wholememory_unique_id_t& uid, int rank, int size) : rank_(rank), size_(size)
{
  WM_EXPECT(size >= 1 && rank >= 0 && rank < size, WHOLEMEMORY_INVALID_INPUT, "bad rank %d / size %d", rank, size);
  if (size == 1) return;
  sockaddr_un addr;
  socklen_t alen = make_addr(uid, &addr);
  hello expect{};
  fill_secret(uid, expect.secret);
  if (rank == 0) {
    peers_.assign(size, -1);
    listen_fd_ = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    WM_EXPECT(listen_fd_ >= 0, WHOLEMEMORY_SYSTEM_ERROR, "socket(): %s", strerror(errno));
    if (::bind(listen_fd_, (sockaddr*)&addr, alen) != 0)
      WM_THROW(WHOLEMEMORY_SYSTEM_ERROR, "bootstrap bind failed (unique id reused?): %s", strerror(errno));
    WM_EXPECT(::listen(listen_fd_, size) == 0, WHOLEMEMORY_SYSTEM_ERROR, "listen(): %s", strerror(errno));
    int64_t deadline = now_ms() + kConnectTimeoutMs;
    for (int got = 1; got < size;) {
      pollfd pfd{listen_fd_, POLLIN, 0};
      int remaining = (int)(deadline - now_ms());
      if (remaining <= 0 || ::poll(&pfd, 1, remaining) <= 0)
        WM_THROW(WHOLEMEMORY_COMMUNICATION_ERROR, "bootstrap: only %d of %d ranks connected", got, size);
      int s = ::accept4(listen_fd_, nullptr, nullptr, SOCK_CLOEXEC);
      if (s < 0) continue;
      /* A connection that is not one of ours -- another user's process, a wrong or missing secret, a rank number that
       * is out of range or already taken, or silence -- is dropped and the wait goes on; it can neither join (and be
       * handed the memory fds) nor stall communicator creation past the connect deadline. */
      hello h{};
      const bool ok = peer_is_our_user(s) && recv_before(s, &h, sizeof(h), std::min(deadline, now_ms() + 5000)) &&
                      memcmp(h.secret, expect.secret, sizeof(h.secret)) == 0 && h.rank > 0 && h.rank < size && peers_[h.rank] < 0;
      if (!ok) {
        WM_WARN("bootstrap: dropped a connection that did not identify as a rank of this communicator");
        ::close(s);
        continue;
      }
      peers_[h.rank] = s;
      ++got;
    }
    ::close(listen_fd_); /* nobody else may join; frees the abstract name */
    listen_fd_ = -1;
  } else {
    peers_.assign(1, -1);
    int64_t deadline = now_ms() + kConnectTimeoutMs;
    for (;;) {
      int s = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
      WM_EXPECT(s >= 0, WHOLEMEMORY_SYSTEM_ERROR, "socket(): %s", strerror(errno));
      if (::connect(s, (sockaddr*)&addr, alen) == 0) {
        if (!peer_is_our_user(s)) { /* somebody else bound the name first */
          ::close(s);
          WM_THROW(WHOLEMEMORY_COMMUNICATION_ERROR, "bootstrap: the listener for this unique id belongs to another user");
        }
        peers_[0] = s;
        break;
      }
      ::close(s);
      if (now_ms() > deadline)
        WM_THROW(WHOLEMEMORY_COMMUNICATION_ERROR, "bootstrap: rank %d could not reach rank 0: %s", rank, strerror(errno));
      usleep(2000);
    }
    hello me = expect;
    me.rank  = rank;
    send_all(peers_[0], &me, sizeof(me));
  }
  barrier();
  setup_mailbox();
}

bootstrap::~bootstrap()
{
  if (mbox_ != nullptr) {
    if (mbox_->slots != nullptr) ::munmap(mbox_->slots, mbox_->bytes);
    delete mbox_;
  }
  for (int s : peers_)
    if (s >= 0) ::close(s);
  if (listen_fd_ >= 0) ::close(listen_fd_);
}

/* Rank 0 creates an anonymous memory file, every rank maps it; the ranks then agree (over the sockets) that all of them
 * succeeded -- otherwise everybody stays on the socket path. */
void bootstrap::setup_mailbox()
{
  const size_t bytes = sizeof(mailbox::slot) * 2 * (size_t)size_;
  int fd             = -1;
  if (rank_ == 0) {
    fd = ::memfd_create("wgb200-mailbox", MFD_CLOEXEC);
    if (fd >= 0 && ::ftruncate(fd, (off_t)bytes) != 0) {
      ::close(fd);
      fd = -1;
    }
    for (int r = 1; r < size_; ++r) send_fd(peers_[r], fd);
  } else {
    fd = recv_fd(peers_[0]);
  }
  void* map = MAP_FAILED;
  if (fd >= 0) {
    map = ::mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0); /* a fresh memfd reads as zeros: seq = 0 */
    ::close(fd);
  }
  char ok = map != MAP_FAILED ? 1 : 0;
  std::vector<char> all(size_);
  socket_allgather(&ok, all.data(), 1);
  const bool everyone = std::all_of(all.begin(), all.end(), [](char c) { return c == 1; });
  if (!everyone) {
    if (map != MAP_FAILED) ::munmap(map, bytes);
    WM_WARN("bootstrap: shared-memory mailbox unavailable, small collectives stay on the sockets");
    return;
  }
  mbox_        = new mailbox();
  mbox_->slots = static_cast<mailbox::slot*>(map);
  mbox_->bytes = bytes;
}

/* a peer that died closes its socket: turn an endless spin into an error */
void bootstrap::check_sockets_alive()
{
  for (int s : peers_) {
    if (s < 0) continue;
    pollfd pfd{s, POLLRDHUP, 0};
    if (::poll(&pfd, 1, 0) > 0 && (pfd.revents & (POLLRDHUP | POLLHUP | POLLERR)))
      WM_THROW(WHOLEMEMORY_COMMUNICATION_ERROR, "bootstrap peer closed the connection");
  }
}

void bootstrap::mailbox_allgather(const void* send, void* recv, size_t bytes)
{
  mailbox& m       = *mbox_;
  const uint64_t k = ++m.round;
  mailbox::slot* bank = m.slots + (size_t)(k & 1) * size_;
  memcpy(bank[rank_].data, send, bytes);
  bank[rank_].seq.store(k, std::memory_order_release);
  char* out           = static_cast<char*>(recv);
  const int limit_ms  = recv_timeout_ms();
  int64_t started     = 0;
  for (int r = 0; r < size_; ++r) {
    if (r != rank_) {
      uint32_t spins = 0;
      while (bank[r].seq.load(std::memory_order_acquire) != k) {
        ++spins;
        if (spins < 4096) {
          _mm_pause();
        } else {
          ::sched_yield(); /* ranks may share cores (tests run 3 ranks on a small box) */
          if ((spins & 1023) == 0) {
            check_sockets_alive();
            const int64_t now = now_ms();
            if (started == 0) started = now;
            if (limit_ms >= 0 && now - started > limit_ms)
              WM_THROW(WHOLEMEMORY_COMMUNICATION_ERROR,
                       "bootstrap: no answer from rank %d within %d s (ranks issued different collectives?)", r, limit_ms / 1000);
          }
        }
      }
    }
    memcpy(out + (size_t)r * bytes, bank[r].data, bytes);
  }
}

void bootstrap::socket_allgather(const void* send, void* recv, size_t bytes)
{
  char* out = static_cast<char*>(recv);
  if (rank_ == 0) {
    memmove(out, send, bytes);
    for (int r = 1; r < size_; ++r) recv_all(peers_[r], out + (size_t)r * bytes, bytes);
    for (int r = 1; r < size_; ++r) send_all(peers_[r], out, bytes * size_);
  } else {
    send_all(peers_[0], send, bytes);
    recv_all(peers_[0], out, bytes * size_);
  }
}

void bootstrap::allgather(const void* send, void* recv, size_t bytes)
{
  if (size_ == 1) {
    if (recv != send) memcpy(recv, send, bytes);
    return;
  }
  /* `bytes` is the same on every rank (it is a collective), so every rank takes the same path */
  if (mbox_ != nullptr && bytes <= mailbox::kPayload) {
    if (recv == send) { /* in-place: the payload would be overwritten while slot 0 is copied out */
      std::vector<char> mine(static_cast<const char*>(send), static_cast<const char*>(send) + bytes);
      mailbox_allgather(mine.data(), recv, bytes);
    } else {
      mailbox_allgather(send, recv, bytes);
    }
    return;
  }
  socket_allgather(send, recv, bytes);
}

void bootstrap::barrier()
{
  char token = 0;
  std::vector<char> all(size_);
  allgather(&token, all.data(), 1);
}

void bootstrap::broadcast(void* buf, size_t bytes, int root)
{
  if (size_ == 1) return;
  std::vector<char> all(bytes * size_);
  allgather(buf, all.data(), bytes);
  memcpy(buf, all.data() + (size_t)root * bytes, bytes);
}

void bootstrap::alltoall(const void* send, void* recv, size_t bytes)
{
  if (size_ == 1) {
    if (recv != send) memcpy(recv, send, bytes);
    return;
  }
  /* small payloads only (per-rank counts): gather the whole matrix, keep our column */
  std::vector<char> all(bytes * size_ * size_);
  allgather(send, all.data(), bytes * size_);
  char* out = static_cast<char*>(recv);
  for (int src = 0; src < size_; ++src)
    memcpy(out + (size_t)src * bytes, all.data() + ((size_t)src * size_ + rank_) * bytes, bytes);
}

std::vector<int> bootstrap::allgather_fds(int my_fd)
{
  std::vector<int> fds(size_, -1);
  if (size_ == 1) {
    fds[0] = my_fd >= 0 ? ::dup(my_fd) : -1;
    return fds;
  }
  if (rank_ == 0) {
    fds[0] = my_fd >= 0 ? ::dup(my_fd) : -1;
    for (int r = 1; r < size_; ++r) fds[r] = recv_fd(peers_[r]);
    for (int r = 1; r < size_; ++r)
      for (int src = 0; src < size_; ++src) send_fd(peers_[r], fds[src]);
  } else {
    send_fd(peers_[0], my_fd);
    for (int src = 0; src < size_; ++src) fds[src] = recv_fd(peers_[0]);
  }
  return fds;
}

}  // namespace wm
