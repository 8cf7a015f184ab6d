// query.cuh — placeholder, filled in below.
#pragma once
