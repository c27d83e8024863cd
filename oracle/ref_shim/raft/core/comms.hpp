#pragma once
namespace raft {
namespace comms {
enum class status_t { SUCCESS, ERROR, ABORT };
}
}  // namespace raft
