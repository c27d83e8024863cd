/*
 * Stress / correctness test of the control-plane bootstrap (wholegraph_b200/csrc/bootstrap.cpp): N forked ranks run
 * thousands of allgather / alltoall / broadcast / barrier rounds with payload sizes on both sides of the shared-memory
 * mailbox limit (so mailbox and socket rounds interleave), every byte checked; then the time of a barrier is printed.
 * Also: an intruder that connects to the listener with a wrong secret (or says nothing) must neither join nor stall it.
 *   bootstrap_stress <ranks> <rounds>
 */
#include <sys/socket.h>
#include <sys/un.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "wm_internal.hpp"

static unsigned char pattern(int round, int rank, size_t i) { return (unsigned char)(round * 131 + rank * 17 + i * 7 + 3); }

static int rank_body(const wholememory_unique_id_t& uid, int rank, int size, int rounds)
{
  try {
    wm::bootstrap b(uid, rank, size);
    const size_t sizes[] = {1, 4, 8, 128, 1016, 1017, 3000, 64, 16};
    int bad = 0;
    for (int k = 0; k < rounds; ++k) {
      size_t n = sizes[k % (sizeof(sizes) / sizeof(sizes[0]))];
      std::vector<unsigned char> mine(n), all(n * size);
      for (size_t i = 0; i < n; ++i) mine[i] = pattern(k, rank, i);
      b.allgather(mine.data(), all.data(), n);
      for (int r = 0; r < size; ++r)
        for (size_t i = 0; i < n; ++i) bad += all[r * n + i] != pattern(k, r, i);
      if (k % 7 == 0) { /* alltoall of 8-byte counts: recv[src] = value src addressed to me */
        std::vector<int64_t> snd(size), rcv(size);
        for (int d = 0; d < size; ++d) snd[d] = (int64_t)k * 1000003 + rank * 1009 + d;
        b.alltoall(snd.data(), rcv.data(), sizeof(int64_t));
        for (int s = 0; s < size; ++s) bad += rcv[s] != (int64_t)k * 1000003 + s * 1009 + rank;
      }
      if (k % 11 == 0) {
        int64_t v = rank == k % size ? 77 + k : -1;
        b.broadcast(&v, sizeof(v), k % size);
        bad += v != 77 + k;
      }
      if (k % 13 == 0) b.barrier();
    }
    b.barrier();
    auto t0 = std::chrono::steady_clock::now();
    const int reps = 2000;
    for (int i = 0; i < reps; ++i) b.barrier();
    double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
    if (rank == 0) {
      printf("barrier: %.2f us per round at %d ranks\n", us, size);
      fflush(stdout);
    }
    if (bad) fprintf(stderr, "rank %d: %d mismatches\n", rank, bad);
    return bad ? 1 : 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "rank %d: exception %s\n", rank, e.what());
    return 2;
  }
}

/* connects to rank 0's listener like a stranger would (the socket name is public in /proc/net/unix) */
static void intruder(const wholememory_unique_id_t& uid, int how)
{
  sockaddr_un addr;
  memset(&addr, 0, sizeof(addr));
  addr.sun_family = AF_UNIX;
  char* p = addr.sun_path + 1;
  int n   = snprintf(p, sizeof(addr.sun_path) - 1, "wgb200-");
  for (int i = 0; i < 12; ++i) n += snprintf(p + n, sizeof(addr.sun_path) - 1 - n, "%02x", (unsigned char)uid.internal[i]);
  socklen_t alen = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
  for (int tries = 0; tries < 2000; ++tries) {
    int s = socket(AF_UNIX, SOCK_STREAM, 0);
    if (connect(s, (sockaddr*)&addr, alen) == 0) {
      if (how == 0) { /* plausible rank number, wrong secret */
        struct {
          int32_t rank;
          unsigned char secret[16];
        } h{1, {0}};
        (void)!write(s, &h, sizeof(h));
      }
      /* how == 1: say nothing, hold the connection */
      sleep(8);
      close(s);
      return;
    }
    close(s);
    usleep(1000);
  }
}

int main(int argc, char** argv)
{
  int size   = argc > 1 ? atoi(argv[1]) : 4;
  int rounds = argc > 2 ? atoi(argv[2]) : 3000;
  bool with_intruders = argc > 3 && atoi(argv[3]) != 0;
  wholememory_unique_id_t uid;
  if (wholememory_create_unique_id(&uid) != WHOLEMEMORY_SUCCESS) return 3;
  std::vector<pid_t> kids;
  if (with_intruders)
    for (int how = 0; how < 2; ++how) {
      pid_t pid = fork();
      if (pid == 0) {
        intruder(uid, how);
        _exit(0);
      }
      kids.push_back(pid);
    }
  if (with_intruders) usleep(20000); /* let the intruders start polling for the listener first */
  std::vector<pid_t> ranks;
  for (int r = 0; r < size; ++r) {
    pid_t pid = fork();
    if (pid == 0) {
      if (with_intruders && r != 0) usleep(300000); /* the intruders reach the listener before the real ranks */
      _exit(rank_body(uid, r, size, rounds));
    }
    ranks.push_back(pid);
  }
  int failures = 0;
  for (pid_t p : ranks) {
    int st = 0;
    waitpid(p, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) ++failures;
  }
  for (pid_t p : kids) {
    int st = 0;
    waitpid(p, &st, 0);
  }
  printf("%d failed ranks\n", failures);
  return failures ? 1 : 0;
}
