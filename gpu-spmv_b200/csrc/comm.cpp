// comm.cpp -- SocketComm / ThreadComm (see comm.hpp).  Plain POSIX, no CUDA.
#include "comm.hpp"

#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/types.h>
#include <sys/un.h>
#include <unistd.h>

namespace spmv {
namespace b200 {
namespace {

bool send_all(int fd, const void* buf, size_t n) {
    const char* p = static_cast<const char*>(buf);
    while (n > 0) {
        const ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
        if (k < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        p += k;
        n -= static_cast<size_t>(k);
    }
    return true;
}

bool recv_all(int fd, void* buf, size_t n) {
    char* p = static_cast<char*>(buf);
    while (n > 0) {
        const ssize_t k = ::recv(fd, p, n, 0);
        if (k < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        if (k == 0) return false;  // peer closed
        p += k;
        n -= static_cast<size_t>(k);
    }
    return true;
}

// one descriptor as SCM_RIGHTS ancillary data next to a 1-byte payload
bool send_fd(int sock, int fd) {
    char byte = 'F';
    iovec iov{&byte, 1};
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
    std::memset(ctrl, 0, sizeof(ctrl));
    msghdr msg{};
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    cmsghdr* c = CMSG_FIRSTHDR(&msg);
    c->cmsg_level = SOL_SOCKET;
    c->cmsg_type = SCM_RIGHTS;
    c->cmsg_len = CMSG_LEN(sizeof(int));
    std::memcpy(CMSG_DATA(c), &fd, sizeof(int));
    for (;;) {
        const ssize_t k = ::sendmsg(sock, &msg, MSG_NOSIGNAL);
        if (k == 1) return true;
        if (k < 0 && errno == EINTR) continue;
        return false;
    }
}

int recv_fd(int sock) {
    char byte = 0;
    iovec iov{&byte, 1};
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
    std::memset(ctrl, 0, sizeof(ctrl));
    msghdr msg{};
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    for (;;) {
        const ssize_t k = ::recvmsg(sock, &msg, MSG_CMSG_CLOEXEC);
        if (k < 0 && errno == EINTR) continue;
        if (k != 1) return -1;
        break;
    }
    for (cmsghdr* c = CMSG_FIRSTHDR(&msg); c; c = CMSG_NXTHDR(&msg, c)) {
        if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) {
            int fd = -1;
            std::memcpy(&fd, CMSG_DATA(c), sizeof(int));
            return fd;
        }
    }
    return -1;
}

void set_timeouts(int fd, int seconds) {
    timeval tv{seconds, 0};
    setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
}

// abstract namespace (no file to clean up): sun_path[0] == 0
socklen_t abstract_address(const char* session, sockaddr_un* addr) {
    std::memset(addr, 0, sizeof(*addr));
    addr->sun_family = AF_UNIX;
    std::string name = std::string("spmv_b200.") + (session ? session : "default");
    if (name.size() > sizeof(addr->sun_path) - 2) name.resize(sizeof(addr->sun_path) - 2);
    std::memcpy(addr->sun_path + 1, name.data(), name.size());
    return static_cast<socklen_t>(offsetof(sockaddr_un, sun_path) + 1 + name.size());
}

class SocketComm final : public Comm {
public:
    SocketComm(int rank, int world) { rank_ = rank; world_ = world; peers_.assign(world, -1); }
    ~SocketComm() override {
        for (int fd : peers_) if (fd >= 0) ::close(fd);
        if (hub_ >= 0) ::close(hub_);
        if (listen_ >= 0) ::close(listen_);
    }

    int connect_all(const char* session, int timeout_s) {
        sockaddr_un addr;
        const socklen_t len = abstract_address(session, &addr);
        if (world_ == 1) return 0;
        if (rank_ == 0) {
            listen_ = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
            if (listen_ < 0) return -1;
            if (::bind(listen_, reinterpret_cast<sockaddr*>(&addr), len) != 0) return -1;
            if (::listen(listen_, world_) != 0) return -1;
            set_timeouts(listen_, timeout_s);
            for (int k = 1; k < world_; ++k) {
                const int fd = ::accept4(listen_, nullptr, nullptr, SOCK_CLOEXEC);
                if (fd < 0) return -1;
                set_timeouts(fd, timeout_s);
                int32_t who = -1;
                if (!recv_all(fd, &who, sizeof(who)) || who <= 0 || who >= world_ || peers_[who] >= 0) {
                    ::close(fd);
                    return -1;
                }
                peers_[who] = fd;
            }
            return 0;
        }
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(timeout_s);
        for (;;) {  // rank 0 may not be listening yet
            hub_ = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
            if (hub_ < 0) return -1;
            if (::connect(hub_, reinterpret_cast<sockaddr*>(&addr), len) == 0) break;
            ::close(hub_);
            hub_ = -1;
            if (std::chrono::steady_clock::now() > deadline) return -1;
            std::this_thread::sleep_for(std::chrono::milliseconds(20));
        }
        set_timeouts(hub_, timeout_s);
        const int32_t who = rank_;
        return send_all(hub_, &who, sizeof(who)) ? 0 : -1;
    }

    int allgather(const void* send, void* recv, size_t bytes) override {
        char* out = static_cast<char*>(recv);
        if (world_ == 1) {
            std::memcpy(out, send, bytes);
            return 0;
        }
        if (rank_ == 0) {
            std::memcpy(out, send, bytes);
            for (int p = 1; p < world_; ++p)
                if (!recv_all(peers_[p], out + p * bytes, bytes)) return -1;
            for (int p = 1; p < world_; ++p)
                if (!send_all(peers_[p], out, bytes * world_)) return -1;
            return 0;
        }
        if (!send_all(hub_, send, bytes)) return -1;
        return recv_all(hub_, out, bytes * world_) ? 0 : -1;
    }

    int barrier() override {
        char mine = 1;
        std::vector<char> all(world_);
        return allgather(&mine, all.data(), 1);
    }

    int allgather_fds(int my_fd, int* fds_out) override {
        for (int p = 0; p < world_; ++p) fds_out[p] = -1;
        if (world_ == 1) {
            fds_out[0] = ::dup(my_fd);
            return fds_out[0] >= 0 ? 0 : -1;
        }
        int rc = 0;
        if (rank_ == 0) {
            fds_out[0] = ::dup(my_fd);
            if (fds_out[0] < 0) rc = -1;
            for (int p = 1; p < world_; ++p) {
                fds_out[p] = recv_fd(peers_[p]);
                if (fds_out[p] < 0) rc = -1;
            }
            for (int p = 1; p < world_ && rc == 0; ++p)
                for (int q = 0; q < world_; ++q)
                    if (!send_fd(peers_[p], fds_out[q])) { rc = -1; break; }
        } else {
            if (!send_fd(hub_, my_fd)) rc = -1;
            for (int q = 0; q < world_ && rc == 0; ++q) {
                fds_out[q] = recv_fd(hub_);
                if (fds_out[q] < 0) rc = -1;
            }
        }
        if (rc != 0)
            for (int p = 0; p < world_; ++p)
                if (fds_out[p] >= 0) { ::close(fds_out[p]); fds_out[p] = -1; }
        return rc;
    }

private:
    int listen_ = -1, hub_ = -1;
    std::vector<int> peers_;  // rank 0: one connection per rank
};

}  // namespace

int comm_create_socket(int rank, int world, const char* session, int timeout_s, Comm** out) {
    if (!out || world < 1 || rank < 0 || rank >= world) return -1;
    SocketComm* c = new SocketComm(rank, world);
    if (c->connect_all(session, timeout_s > 0 ? timeout_s : 120) != 0) {
        delete c;
        return -1;
    }
    *out = c;
    return 0;
}

// ---------------------------------------------------------------------------- threads ----

class ThreadCommGroup {
public:
    explicit ThreadCommGroup(int world) : world_(world), slots_(world, nullptr), fds_(world, -1) {}
    int world_;
    std::mutex mu_;
    std::condition_variable cv_;
    int arrived_ = 0;
    unsigned long long generation_ = 0;
    std::vector<const void*> slots_;
    std::vector<int> fds_;
    std::vector<Comm*> comms_;

    void sync() {  // generation barrier
        std::unique_lock<std::mutex> lock(mu_);
        const unsigned long long gen = generation_;
        if (++arrived_ == world_) {
            arrived_ = 0;
            ++generation_;
            cv_.notify_all();
        } else {
            cv_.wait(lock, [&] { return generation_ != gen; });
        }
    }
};

namespace {

class ThreadComm final : public Comm {
public:
    ThreadComm(ThreadCommGroup* g, int rank) : g_(g) { rank_ = rank; world_ = g->world_; }
    int allgather(const void* send, void* recv, size_t bytes) override {
        g_->slots_[rank_] = send;
        g_->sync();
        for (int p = 0; p < world_; ++p) std::memcpy(static_cast<char*>(recv) + p * bytes, g_->slots_[p], bytes);
        g_->sync();  // nobody reuses its send buffer before everyone has copied
        return 0;
    }
    int barrier() override {
        g_->sync();
        return 0;
    }
    int allgather_fds(int my_fd, int* fds_out) override {
        g_->fds_[rank_] = my_fd;
        g_->sync();
        int rc = 0;
        for (int p = 0; p < world_; ++p) {
            fds_out[p] = ::dup(g_->fds_[p]);
            if (fds_out[p] < 0) rc = -1;
        }
        g_->sync();
        return rc;
    }
private:
    ThreadCommGroup* g_;
};

}  // namespace

ThreadCommGroup* thread_comm_group_create(int world) {
    if (world < 1) return nullptr;
    ThreadCommGroup* g = new ThreadCommGroup(world);
    for (int r = 0; r < world; ++r) g->comms_.push_back(new ThreadComm(g, r));
    return g;
}
Comm* thread_comm_get(ThreadCommGroup* g, int rank) { return (g && rank >= 0 && rank < g->world_) ? g->comms_[rank] : nullptr; }
void thread_comm_group_destroy(ThreadCommGroup* g) {
    if (!g) return;
    for (Comm* c : g->comms_) delete c;
    delete g;
}

}  // namespace b200
}  // namespace spmv
